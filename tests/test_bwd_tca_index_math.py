"""CPU model of the index arithmetic of csrc/corr1d_bwd_tca.cu (no GPU, no library call): the g slice layouts the TMA
producer writes, the diagonal reads of the A builders (mode 1: rotated column order + register rotation), and the
box wait / release schedule.  It restates the kernel's formulas line by line, so a change of one side without the
other fails here before it reaches a GPU."""
import numpy as np
import pytest

KTM, KKC, PITCH0, PITCH1, GROUPS, ROW0 = 128, 32, 128, 136, 2, 4


def fill_args(P):
    a = {}
    r = (P - 1) // 2
    oo0, oo1 = -r, -(P - 1 - r)
    a["r"] = r
    a["delta"] = [oo0 % 4, oo1 % 4]
    a["oo"] = [oo0 - a["delta"][0], oo1 - a["delta"][1]]
    a["koff0"] = (KTM + KKC - 2 + a["delta"][0]) // KKC
    a["c1"] = P - 1 + a["delta"][1]
    a["e0"] = (r - 3) % 4
    a["NKC"] = -(-(KTM + P - 1 + max(a["delta"])) // KKC)
    a["n_gboxes"] = -(-P // 32)
    return a


def load_slices(g, x0, P, W, a):
    """What the TMA producer leaves in shared memory (zero fill outside the tensor)."""
    nb = a["n_gboxes"]
    # ROW0 zero rows in front of plane 0 and ROW0 after the last box (never written by the TMA)
    s0 = np.zeros((32 * nb + 2 * ROW0, PITCH0), np.float64)
    s1 = np.zeros((32 * nb + 2 * ROW0, PITCH1), np.float64)
    for p in range(32 * nb):
        for i in range(PITCH0):
            w = x0 + i
            if p < P and 0 <= w < W:
                s0[p + ROW0, i] = g[p, w]
    for pg in range(0, 32 * nb, 4):
        start = x0 - (pg + 3) + a["r"] - a["e0"]
        assert start % 4 == 0, "TMA needs a 16-byte aligned inner start"
        for p in range(pg, pg + 4):
            for i in range(PITCH1):
                w = start + i
                if p < P and 0 <= w < W:
                    s1[p + ROW0, i] = g[p, w]
    return s0, s1


def build_chunk(mode, s, xl, k, P, a):
    """The 32 band columns one builder thread (TMEM lane xl) produces for chunk k, plus the planes it read."""
    w = np.zeros(32)
    planes = []

    def row(p):   # the kernel's branch-free clamp: umin(unsigned(p + 1), P + 1) + ROW0 - 1
        q = (p + 1) & 0xFFFFFFFF
        if 0 <= p < P:
            planes.append(p)
        return min(q, P + 1) + ROW0 - 1

    if mode == 0:
        pu = KKC * k - a["delta"][0] - xl
        for t in range(32):
            w[t] = s[row(pu + t), xl]
    else:
        bu = [xl - ((u + xl) & 3) for u in range(4)]
        cu = [xl + 3 - ((a["c1"] - u) & 3) + a["e0"] for u in range(4)]
        pk = a["c1"] - KKC * k
        srot = xl & 3
        for gg in range(8):
            v = [0.0] * 4
            for u in range(4):
                assert 0 <= cu[u] < PITCH1
                v[u] = s[row(bu[u] + pk - 4 * gg), cu[u]]
            if srot & 1:
                v = [v[3], v[0], v[1], v[2]]
            if srot & 2:
                v = [v[2], v[3], v[0], v[1]]
            w[4 * gg:4 * gg + 4] = v
    return w, planes


@pytest.mark.parametrize("P", [1, 7, 17, 40, 192, 193])
def test_gd_rows_match_the_definition(P):
    rng = np.random.default_rng(P)
    W = 256
    g = rng.standard_normal((P, W))
    a = fill_args(P)
    for x0 in (0, 128):
        s0, s1 = load_slices(g, x0, P, W, a)
        for mode, s in ((0, s0), (1, s1)):
            for xl in list(range(0, 128, 7)) + [127]:
                for k in range(a["NKC"]):
                    got, _ = build_chunk(mode, s, xl, k, P, a)
                    for jj in range(32):
                        j = KKC * k + jj
                        if mode == 0:
                            p, wcol = j - a["delta"][0] - xl, x0 + xl
                        else:
                            p, wcol = xl + a["c1"] - j, x0 + a["oo"][1] + j
                        want = g[p, wcol] if (0 <= p < P and 0 <= wcol < W) else 0.0
                        assert got[jj] == want, (P, mode, x0, xl, k, jj)


@pytest.mark.parametrize("P", [1, 17, 40, 100, 192, 193, 250])
def test_mode1_reads_are_bank_conflict_free_and_boxes_are_scheduled_safely(P):
    a = fill_args(P)
    nb, NKC = a["n_gboxes"], a["NKC"]
    # bank conflicts: one LDS instruction = fixed (gg, u), lanes xl = 32q .. 32q+31
    for q in range(4):
        for u in range(4):
            banks = set()
            for lane in range(32):
                xl = 32 * q + lane
                p = xl - ((u + xl) & 3) + a["c1"]
                col = xl + 3 - ((a["c1"] - u) & 3) + a["e0"]
                banks.add((p * PITCH1 + col) % 32)
            assert len(banks) == 32
    # box schedule, per builder group: every plane read by chunk k lies in a box that was waited for and not yet released
    for mode in (0, 1):
        for first in range(GROUPS):           # which chunks of the tile this group visits (k = first, first+GROUPS, ...)
            ready, rel = 0, 0
            for k in range(first, NKC, GROUPS):
                if mode == 0:
                    need = min(k + 1, nb)
                else:
                    lowp = a["c1"] - KKC * k - (KKC - 1)
                    lowb = min(max(lowp, 0) >> 5 if lowp > 0 else 0, nb - 1)
                    need = nb - lowb
                ready = max(ready, need)
                held = set(range(rel, ready)) if mode == 0 else set(nb - 1 - o for o in range(rel, ready))
                for xl in (0, 31, 64, 127):
                    _, planes = build_chunk(mode, np.zeros((32 * nb + 2 * ROW0, PITCH1)), xl, k, P, a)
                    for p in planes:
                        assert (p >> 5) in held, (P, mode, k, xl, p, sorted(held))
                last_visit = k + GROUPS >= NKC
                while rel < nb:
                    kl = rel + a["koff0"] if mode == 0 else (a["c1"] + KTM - 1 - 32 * (nb - 1 - rel)) >> 5
                    kl = min(kl, NKC - 1)
                    if not (kl <= k + GROUPS - 1 or last_visit):
                        break
                    rel += 1
            assert rel == nb and ready <= nb   # everything handed back by the group's last visit
