"""CPU check of the row-staged K1 kernel's ADDRESSING scheme (csrc/va_small_kernels.cu::preprocess_rows_kernel): a
Python restatement of its staging (one 16-byte-granular bulk copy per (row, plane), two buffers) and of its convert
mapping (thread t owns chunks t, t+224, ... of every row; per-plane byte shift and flip), run on raw bytes against a
plain crop/flip.  It pins the index arithmetic -- shifts, plane offsets, flip, chunk -> (pixel, channel group), output
offsets, no read outside the store -- without a GPU; the GPU parity tests (tests/test_gpu_preprocess.py) check the
kernel itself bit-exactly against the oracle."""
import numpy as np
import pytest

CROP, THREADS = 224, 224


def emulate_rows_kernel(images, image_bytes, img_w, IMG_C, PLANES, C_PAD, RB, table, n, grid):
    NCH = PLANES * IMG_C
    JN, CPC = min(NCH, 8), C_PAD // 8
    XSTEP, ROWB = CROP // CPC, CROP * IMG_C
    COPYB = (ROWB + 30) // 16 * 16
    ROWSZ = PLANES * COPYB + (PLANES + 7) // 8 * 32
    BUFB, NUNITS, IPS = RB * ROWSZ, RB * PLANES, CROP // RB
    ploff = lambda pl: pl * COPYB + (pl >> 3) * 32
    out = np.full(n * CROP * CROP * C_PAD, -1, np.int32)
    raw0 = np.full(2 * BUFB, 255, np.uint8)
    rowoff0, flip0 = np.zeros(2 * NUNITS, np.int64), np.zeros(2 * PLANES, np.int64)

    def stage(item, buf):
        for tid in range(NUNITS):
            r, pl = divmod(tid, PLANES)
            snip = item // IPS
            y0 = (item - snip * IPS) * RB
            iid, ci, cj, fl = (int(v) for v in table[snip, pl])
            off = iid * image_bytes + ((ci + y0 + r) * img_w + cj) * IMG_C
            sh = off & 15
            nbytes = (sh + ROWB + 15) & ~15
            assert nbytes <= COPYB and off - sh + nbytes <= images.size          # never outside the store
            dst = r * ROWSZ + ploff(pl)
            rowoff0[buf * NUNITS + tid] = dst + sh + ((CROP - 1) * IMG_C if fl else 0)
            if r == 0:
                flip0[buf * PLANES + pl] = 1 if fl else 0
            raw0[buf * BUFB + dst: buf * BUFB + dst + nbytes] = images[off - sh: off - sh + nbytes]

    n_items = n * IPS
    for blk in range(grid):
        item, it = blk, 0
        if item < n_items:
            stage(item, 0)
        while item < n_items:
            buf = it & 1
            if item + grid < n_items:
                stage(item + grid, buf ^ 1)
            raw = raw0[buf * BUFB:(buf + 1) * BUFB]
            snip = item // IPS
            y0 = (item - snip * IPS) * RB
            chunk0 = (snip * CROP + y0) * CROP * C_PAD // 8
            for tid in range(THREADS):
                g, x0 = tid % CPC, tid // CPC
                for r in range(RB):
                    for i in range(CPC):
                        vals = [-2] * 8                                              # -2 = zero padding channel
                        for j in range(JN):
                            c = g * 8 + j
                            if c < NCH:
                                pl = c // IMG_C
                                step = -IMG_C if flip0[buf * PLANES + pl] else IMG_C
                                a = rowoff0[buf * NUNITS + r * PLANES + pl] + x0 * step + (c - pl * IMG_C)
                                vals[j] = int(raw[a + i * XSTEP * step])
                        ch = chunk0 + tid + (r * CPC + i) * THREADS
                        out[ch * 8: ch * 8 + 8] = vals
            item += grid
            it += 1
    return out.reshape(n, CROP, CROP, C_PAD)


def crop_flip(images, image_bytes, H, W, IMG_C, PLANES, C_PAD, table, n):
    out = np.full((n, CROP, CROP, C_PAD), -2, np.int32)
    for s in range(n):
        for pl in range(PLANES):
            iid, ci, cj, fl = (int(v) for v in table[s, pl])
            img = images[iid * image_bytes: iid * image_bytes + H * W * IMG_C].reshape(H, W, IMG_C)
            c = img[ci:ci + CROP, cj:cj + CROP]
            out[s, :, :, pl * IMG_C:(pl + 1) * IMG_C] = c[:, ::-1] if fl else c
    return out


@pytest.mark.parametrize("H,W,IMG_C,PLANES,C_PAD,RB", [(256, 340, 1, 20, 32, 4), (240, 320, 3, 1, 16, 7)])
def test_rows_kernel_addressing(H, W, IMG_C, PLANES, C_PAD, RB):
    rng = np.random.default_rng(H)
    n_img, n = 3, 1
    image_bytes = H * W * IMG_C
    assert image_bytes % 16 == 0                                  # the kernel's dispatch condition
    images = rng.integers(0, 256, n_img * image_bytes, dtype=np.uint8)
    table = np.stack([rng.integers(0, n_img, (n, PLANES)), rng.integers(0, H - CROP + 1, (n, PLANES)),
                      rng.integers(0, W - CROP + 1, (n, PLANES)), rng.integers(0, 2, (n, PLANES))], -1)
    table[0, 0] = [n_img - 1, H - CROP, W - CROP, 1]              # the last bytes of the store, flipped
    got = emulate_rows_kernel(images, image_bytes, W, IMG_C, PLANES, C_PAD, RB, table, n, grid=3)
    assert np.array_equal(got, crop_flip(images, image_bytes, H, W, IMG_C, PLANES, C_PAD, table, n))
