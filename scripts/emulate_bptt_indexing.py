"""Design check for rt_bptt.cuh (no GPU needed): re-states the kernel's slicing, staging and mma fragment
index arithmetic in numpy and verifies that the KS x NS partial products sum to dgates . W_hh and that the
cell phase covers every (batch, unit) exactly once."""
import numpy as np
rs = np.random.RandomState(0)
U, B = 512, 32
U4 = 4 * U
KS, NS = U // 64, U // 32
dg = rs.randn(B, U4).astype(np.float64)          # dgates_t rows (one step)
W = rs.randn(U4, U).astype(np.float64)           # W_hh [4U, U]
want = dg @ W                                    # dh carry [B, U]
part = np.zeros((KS, B, U))
DGP = 260
for cta in range(KS * NS):
    i, j = cta % KS, cta // KS
    # staged tile
    dgs = np.zeros(32 * DGP)
    for v in range(32 * 64):
        b, kk = v // 64, (v % 64) * 4
        src = b * U4 + (kk >> 6) * U + 64 * i + (kk & 63)
        dgs[b * DGP + kk: b * DGP + kk + 4] = dg.reshape(-1)[src:src + 4]
    for warp in range(8):
        mi, ni = warp & 1, warp >> 1
        # reconstruct per k-step A (16x8) and B (8x8) from the lanes' registers, official layouts
        D = np.zeros((16, 8))
        for s in range(32):
            A = np.zeros((16, 8)); Bm = np.zeros((8, 8))
            for lane in range(32):
                g, tg = lane >> 2, lane & 3
                arow0 = (16 * mi + g) * DGP + tg
                arow1 = arow0 + 8 * DGP
                a0, a1, a2, a3 = dgs[arow0 + 8 * s], dgs[arow1 + 8 * s], dgs[arow0 + 8 * s + 4], dgs[arow1 + 8 * s + 4]
                A[g, tg] = a0; A[g + 8, tg] = a1; A[g, tg + 4] = a2; A[g + 8, tg + 4] = a3
                ncol = 32 * j + 8 * ni + g
                for h2 in range(2):
                    kk = 8 * s + tg + 4 * h2
                    krow = (kk >> 6) * U + 64 * i + (kk & 63)
                    Bm[tg + 4 * h2, g] = W[krow, ncol]
            D += A @ Bm
        for lane in range(32):
            g, tg = lane >> 2, lane & 3
            b0 = 16 * mi + g
            n0 = 32 * j + 8 * ni + 2 * tg
            part[i, b0, n0] = D[g, 2 * tg]; part[i, b0, n0 + 1] = D[g, 2 * tg + 1]
            part[i, b0 + 8, n0] = D[g + 8, 2 * tg]; part[i, b0 + 8, n0 + 1] = D[g + 8, 2 * tg + 1]
got = part.sum(0)
print("max err", np.abs(got - want).max())
# cell ownership covers every (b, u) exactly once
cnt = np.zeros((B, U), int)
upc = 32 // KS
for cta in range(KS * NS):
    i, j = cta % KS, cta // KS
    for tid in range(256):
        if tid < 32 * upc:
            cnt[tid & 31, 32 * j + upc * i + (tid >> 5)] += 1
print("cell coverage ok:", (cnt == 1).all())
