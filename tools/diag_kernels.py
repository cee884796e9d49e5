#!/usr/bin/env python
"""GPU diagnostics for the tcgen05 layer kernel: each case runs in its own subprocess (a device trap in one case
must not poison the rest), compares against torch fp32 on bf16-rounded operands, and prints an error summary
that localises descriptor/layout mistakes.  Usage: python tools/diag_kernels.py all | case <name>
"""
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# name -> kind, params
CASES = {
    # plain GEMMs (H=W=1): n, in, out, bn
    "fc_128x64x64": ("fc", dict(n=128, fin=64, fout=64, bn=64)),
    "fc_128x256x64": ("fc", dict(n=128, fin=256, fout=64, bn=64)),
    "fc_200x512x256_bn64": ("fc", dict(n=200, fin=512, fout=256, bn=64)),
    "fc_200x512x256_bn128": ("fc", dict(n=200, fin=512, fout=256, bn=128)),
    "fc_200x512x256_bn256": ("fc", dict(n=200, fin=512, fout=256, bn=256)),
    "fc_f32out": ("fc", dict(n=70, fin=1024, fout=256, bn=64, f32=True)),
    "fc_big": ("fc", dict(n=250, fin=25088, fout=4096, bn=0)),
    # convs: n, H, cin, cin_pad, cout, pool, bn, r
    "conv16_c64": ("conv", dict(n=1, H=16, cin=64, cin_pad=64, cout=64, pool=0, bn=64, r=1)),
    "conv16_c64_pool": ("conv", dict(n=2, H=16, cin=64, cin_pad=64, cout=64, pool=1, bn=64, r=1)),
    "conv32_c128_bn128": ("conv", dict(n=2, H=32, cin=128, cin_pad=128, cout=128, pool=1, bn=128, r=1)),
    "conv32_c128_bn256": ("conv", dict(n=2, H=32, cin=128, cin_pad=128, cout=256, pool=0, bn=256, r=1)),
    "conv56_n3": ("conv", dict(n=3, H=56, cin=128, cin_pad=128, cout=256, pool=1, bn=0, r=1)),
    "conv28_n5": ("conv", dict(n=5, H=28, cin=256, cin_pad=256, cout=512, pool=1, bn=0, r=1)),
    "conv14_n33": ("conv", dict(n=33, H=14, cin=512, cin_pad=512, cout=512, pool=1, bn=0, r=1)),
    "conv_ck16": ("conv", dict(n=2, H=32, cin=3, cin_pad=16, cout=64, pool=0, bn=64, r=1)),
    "conv_ck32": ("conv", dict(n=2, H=32, cin=20, cin_pad=32, cout=64, pool=0, bn=64, r=1)),
    "conv_r3_bn64": ("conv", dict(n=2, H=32, cin=64, cin_pad=64, cout=64, pool=1, bn=64, r=3)),
    "conv_r3_bn128": ("conv", dict(n=2, H=32, cin=128, cin_pad=128, cout=128, pool=0, bn=128, r=3)),
    "conv224_64_pool": ("conv", dict(n=4, H=224, cin=64, cin_pad=64, cout=64, pool=1, bn=64, r=1)),
    "conv224_64_pool_r3": ("conv", dict(n=4, H=224, cin=64, cin_pad=64, cout=64, pool=1, bn=64, r=3)),
    "conv224_ck16": ("conv", dict(n=4, H=224, cin=3, cin_pad=16, cout=64, pool=0, bn=64, r=1)),
    "conv224_ck32": ("conv", dict(n=4, H=224, cin=20, cin_pad=32, cout=64, pool=0, bn=64, r=1)),
}


def _summ(ours, ref, name):
    import torch
    ours = ours.float()
    diff = (ours - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 2e-2 * ref.abs().mean().clamp_min(1e-6)
    bad = diff > tol
    out = dict(case=name, max_abs=float(diff.max()), ref_absmean=float(ref.abs().mean()),
               ours_absmean=float(ours.abs().mean()), frac_bad=float(bad.float().mean()), n=ours.numel(),
               nan=int(torch.isnan(ours).sum()))
    if bad.any():
        idx = bad.nonzero()
        out["first_bad"] = idx[:6].tolist()
        for d in range(bad.dim()):
            other = [i for i in range(bad.dim()) if i != d]
            prof = bad.float().mean(dim=other)
            nz = (prof > 0).nonzero().flatten()
            out[f"bad_dim{d}"] = dict(count=int(nz.numel()), of=int(prof.numel()), first=nz[:12].tolist())
    out["ok"] = bool(not bad.any() and out["nan"] == 0)
    return out


def run_case(name):
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from video_analytics_b200 import ops
    kind, p = CASES[name]
    g = torch.Generator(device="cpu").manual_seed(1234)
    dev = "cuda"
    if kind == "fc":
        x = torch.randn(p["n"], p["fin"], generator=g).to(dev).bfloat16()
        w = (torch.randn(p["fout"], p["fin"], generator=g) / p["fin"] ** 0.5).to(dev)
        b = torch.randn(p["fout"], generator=g).to(dev) * 0.1
        y = ops.linear(x, w, b, relu=True, out_f32=bool(p.get("f32")), force_bn=p["bn"])
        torch.cuda.synchronize()
        ref = torch.relu(x.float() @ w.bfloat16().float().t() + b)
        return _summ(y, ref, name)
    n, H, cin, cin_pad, cout = p["n"], p["H"], p["cin"], p["cin_pad"], p["cout"]
    xc = torch.randn(n, cin, H, H, generator=g).to(dev).bfloat16()
    x = torch.zeros(n, H, H, cin_pad, dtype=torch.bfloat16, device=dev)
    x[..., :cin] = xc.permute(0, 2, 3, 1)
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5).to(dev)
    b = (torch.randn(cout, generator=g) * 0.1).to(dev)
    y = ops.conv2d_nhwc(x, w, b, relu=True, pool=bool(p["pool"]), force_bn=p["bn"], force_r=p["r"])
    torch.cuda.synchronize()
    ref = torch.relu(torch.nn.functional.conv2d(xc.float(), w.bfloat16().float(), b, padding=1))
    if p["pool"]:
        ref = torch.nn.functional.max_pool2d(ref, 2, 2)
    ref = ref.permute(0, 2, 3, 1).contiguous()
    return _summ(y, ref, name)


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "case":
        try:
            res = run_case(sys.argv[2])
        except Exception as e:  # noqa
            res = dict(case=sys.argv[2], ok=False, error=f"{type(e).__name__}: {e}"[:600])
        print("RESULT " + json.dumps(res), flush=True)
        return
    names = list(CASES) if len(sys.argv) < 3 else sys.argv[2:]
    nbad = 0
    for nm in names:
        t0 = time.time()
        try:
            pr = subprocess.run([sys.executable, os.path.abspath(__file__), "case", nm], capture_output=True, text=True,
                                timeout=240)
            lines = [l for l in pr.stdout.splitlines() if l.startswith("RESULT ")]
            tail = (pr.stdout[-600:] + pr.stderr[-900:]) if not lines else ""
            print(f"[{nm}] rc={pr.returncode} {time.time()-t0:.1f}s {lines[-1] if lines else 'NO RESULT'}")
            if lines and not json.loads(lines[-1][7:]).get("ok"):
                nbad += 1
                print("   stdout-tail:", pr.stdout[-400:].replace("\n", " | "))
                print("   stderr-tail:", pr.stderr[-400:].replace("\n", " | "))
            if tail:
                nbad += 1
                print("   tail:", tail.replace("\n", " | "))
        except subprocess.TimeoutExpired:
            nbad += 1
            print(f"[{nm}] TIMEOUT")
        sys.stdout.flush()
    print(f"DIAG DONE bad={nbad} of {len(names)}")


if __name__ == "__main__":
    main()
