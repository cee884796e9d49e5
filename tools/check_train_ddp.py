"""Data-parallel training check (run under torchrun, 2+ ranks): the all-reduced, 1/world-scaled gradient arena of a
batch split across ranks equals the single-GPU gradient of the whole batch, and parameters stay identical on all ranks
after the update.  Prints one line 'DDP_OK ...' on rank 0."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_analytics_b200.distributed import init_from_env  # noqa: E402
from video_analytics_b200.spatialModel import build_spatial_torch_model  # noqa: E402
from video_analytics_b200.training import StreamTrainer  # noqa: E402


def main():
    rank, world, local = init_from_env()
    torch.cuda.set_device(local)
    per = 2
    n = per * world
    g = torch.Generator().manual_seed(11)
    x_all = torch.randn(n, 224, 224, 16, generator=g).bfloat16()
    x_all[..., 3:] = 0
    labels_all = torch.randint(1, 101, (n,), generator=g)
    masks_all = [(torch.rand(n, d, generator=g) >= 0.5).to(torch.uint8) for d in (4096, 4096, 256)]
    sl = slice(rank * per, (rank + 1) * per)
    tr = StreamTrainer(build_spatial_torch_model(101, 256, seed=0), None, lr=0.01, momentum=0.9, c_pad=16,
                       process_group=dist.group.WORLD)
    loss, _, _ = tr.forward_backward(x_all[sl].cuda(), labels_all[sl].cuda(), [m[sl].contiguous().cuda() for m in masks_all])
    dist.all_reduce(tr.flat_grad, op=dist.ReduceOp.SUM)
    avg = tr.flat_grad / world
    loss_avg = loss.clone()
    dist.all_reduce(loss_avg)
    loss_avg /= world
    ok = True
    if rank == 0:
        ref = StreamTrainer(build_spatial_torch_model(101, 256, seed=0), None, lr=0.01, momentum=0.9, c_pad=16)
        loss_ref, _, _ = ref.forward_backward(x_all.cuda(), labels_all.cuda(), [m.cuda() for m in masks_all])
        rel = float((avg - ref.flat_grad).norm() / ref.flat_grad.norm())
        dl = abs(float(loss_avg) - float(loss_ref))
        ok = rel < 1e-4 and dl < 1e-5
        print(f"grad rel diff {rel:.3e}, loss diff {dl:.3e}")
    # the real update path: step on every rank, then compare parameters across ranks
    tr.step(x_all[sl].cuda(), labels_all[sl].cuda(), [m[sl].contiguous().cuda() for m in masks_all])
    mine = tr.flat_param.clone()
    dist.broadcast(mine, src=0)
    same = bool(torch.equal(mine, tr.flat_param))
    flag = torch.tensor([1 if same else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    # replicas built from DIFFERENT seeds per rank must come out identical (StreamTrainer.sync_replicas broadcasts rank 0's)
    tr2 = StreamTrainer(build_spatial_torch_model(101, 256, seed=100 + rank), None, lr=0.01, momentum=0.9, c_pad=16,
                        process_group=dist.group.WORLD, grad_allreduce_dtype="fp32")
    p0 = tr2.flat_param.clone()
    dist.broadcast(p0, src=0)
    synced = torch.tensor([1 if torch.equal(p0, tr2.flat_param) else 0], device="cuda")
    dist.all_reduce(synced, op=dist.ReduceOp.MIN)
    # the bf16 gradient payload (default) against the exact fp32 payload: same data, same start -> the parameter UPDATE differs
    # by bf16 rounding of the summed gradients only
    tr3 = StreamTrainer(build_spatial_torch_model(101, 256, seed=100), None, lr=0.01, momentum=0.9, c_pad=16,
                        process_group=dist.group.WORLD, grad_allreduce_dtype="bf16")
    before = tr2.flat_param.clone()
    args = (x_all[sl].cuda(), labels_all[sl].cuda(), [m[sl].contiguous().cuda() for m in masks_all])
    tr2.step(*args)
    tr3.step(*args)
    d32, d16 = tr2.flat_param - before, tr3.flat_param - before
    rel16 = float((d16 - d32).norm() / d32.norm())
    # deferred update (the all-reduce of step k waits at the start of step k+1, SM reservation for the collective): same
    # updates as the immediate schedule, up to the summation order of the weight-gradient atomics
    del tr2, tr3
    torch.cuda.empty_cache()
    trA = StreamTrainer(build_spatial_torch_model(101, 256, seed=7), None, lr=0.01, momentum=0.9, c_pad=16,
                        process_group=dist.group.WORLD, defer_update=False)
    trB = StreamTrainer(build_spatial_torch_model(101, 256, seed=7), None, lr=0.01, momentum=0.9, c_pad=16,
                        process_group=dist.group.WORLD, defer_update=True)
    start = trA.flat_param.clone()
    for _ in range(2):
        trA.step(*args)
        trB.step(*args)
    pending = trB._deferred is not None
    trB.flush()
    moved = float((trA.flat_param - start).norm())
    rel_defer = float((trA.flat_param - trB.flat_param).norm()) / moved
    if rank == 0:
        print(f"deferred vs immediate update: rel diff of the 2-step parameter change {rel_defer:.3e} (pending before flush: {pending})")
        ok = ok and pending and rel_defer < 1e-4
    # the library's own all-reduce kernel (va_allreduce_bf16) over the symmetric gradient arena: NVSwitch multicast form and
    # peer-pointer form, against the fp32 sum of the ranks' bf16-rounded gradients rounded once to bf16 (what both compute)
    import os
    del trA, trB
    torch.cuda.empty_cache()
    own_ok = True
    for mc in ("1", "0"):
        os.environ["VA_ALLREDUCE_MULTICAST"] = mc
        trV = StreamTrainer(build_spatial_torch_model(101, 256, seed=7), None, lr=0.01, momentum=0.9, c_pad=16,
                            process_group=dist.group.WORLD, allreduce_impl="va", defer_update=False)
        trV.forward_backward(*args)
        ref = trV.flat_grad.bfloat16().float()
        dist.all_reduce(ref, op=dist.ReduceOp.SUM)
        ref = ref.bfloat16()
        trV._allreduce_async(0, trV.flat_grad.numel()).wait()
        torch.cuda.synchronize()
        if trV._symm_use_mc:       # in-switch reduction: its accumulation order / denormal handling is the fabric's -- one bf16 ulp
            d = (trV.flat_grad_bf16.float() - ref.float()).abs()
            same_v = bool((d <= 2.0 ** -7 * ref.float().abs() + 1e-37).all())
            n_diff = int((trV.flat_grad_bf16 != ref).sum())
        else:                      # fp32 sum in rank order, rounded once: exactly the reference
            same_v = torch.equal(trV.flat_grad_bf16, ref)
            n_diff = 0
        nz = int((trV.flat_grad_bf16 != 0).sum())
        fl = torch.tensor([1 if same_v and nz > 0 else 0], device="cuda")
        dist.all_reduce(fl, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"va_allreduce_bf16 ({'multicast' if trV._symm_use_mc else 'peer pointers'}; multicast available: "
                  f"{int(trV._symm.multicast_ptr or 0) != 0}): matches the reference sum on all ranks = {bool(int(fl))} "
                  f"({n_diff} of {nz} non-zero elements differ, within one bf16 ulp)")
        own_ok = own_ok and bool(int(fl))
        del trV, ref
        torch.cuda.empty_cache()
    ok = ok and own_ok if rank == 0 else ok
    if rank == 0:
        good = ok and int(flag) == 1 and int(synced) == 1 and rel16 < 1e-2
        print(("DDP_OK" if good else "DDP_FAIL"), f"world={world} params_identical={bool(int(flag))} replicas_synced={bool(int(synced))} "
              f"bf16_vs_fp32_update_rel={rel16:.3e}")
    dist.destroy_process_group()


main()
