"""CPU model of the round-2 index arithmetic of csrc/corr1d_bwd_tc.cu (no GPU, no library call): the byte-offset
clamp of the gin1 TMEM build, the swizzled raw block + rotated column order + register rotation of the gin2 TMEM build,
and the closed-form box release schedule.  It restates the kernel's formulas line by line, so a change of one side
without the other fails here before it reaches a GPU."""
import numpy as np
import pytest

KTM, KKC, RAW_ROWS = 128, 32, 160
U32 = 1 << 32


def modes(P):
    r = (P - 1) // 2
    oo0, oo1 = -r, -(P - 1 - r)
    d0, d1 = oo0 % 4, oo1 % 4
    nkc = -(-(KTM + P - 1 + max(d0, d1)) // KKC)
    return {"delta": (d0, d1), "oo": (oo0 - d0, oo1 - d1), "NKC": nkc, "n_gboxes": -(-P // 32),
            "koff": (KTM + KKC - 2 + d0) // KKC}


@pytest.mark.parametrize("P", [1, 8, 17, 40, 100, 192, 193])
def test_gin1_offset_clamp_reads_the_right_plane_or_a_zero_row(P):
    """offset = 4 xl + 512 row with row = min(plane + 1, P + 1) in UNSIGNED arithmetic: a plane below 0 must wrap far above
    the limit (4 xl < 512 guarantees it) and land on the zero row P + 1, never below the slice."""
    m = modes(P)
    rng = np.random.default_rng(P)
    g = rng.standard_normal((P, KTM))
    rows = m["n_gboxes"] * 32 + 2
    slice_ = np.zeros((rows, KTM))
    slice_[1:P + 1] = g                        # row 0 and the rows after plane P-1 are zero
    flat = slice_.reshape(-1)
    delta = m["delta"][0]
    for k in range(m["NKC"]):
        for xl in (0, 1, 31, 32, 77, 126, 127):
            pb = KKC * k - delta - xl
            lane_off = (xl * 4) % U32
            lim_off = (lane_off + (P + 1) * KTM * 4) % U32
            for c0 in (0, 16):
                o0 = (lane_off + ((pb + c0 + 1) % U32) * (KTM * 4)) % U32
                for t in range(16):
                    off = min((o0 + t * KTM * 4) % U32, lim_off)
                    assert off % 4 == 0 and 0 <= off // 4 < flat.size
                    got = flat[off // 4]
                    p = pb + c0 + t                       # Gd[xl][32k + c0 + t] = g[p][xl], zero outside [0, P)
                    want = g[p, xl] if 0 <= p < P else 0.0
                    assert got == want, (k, xl, c0, t, p)


def swizzle128_store(raw):
    """Bytes of a [rows][32 floats] TMA box written with CU_TENSOR_MAP_SWIZZLE_128B into a 1024-byte aligned slot:
    the 16-byte chunk index of a row is XORed with (row & 7)."""
    rows = raw.shape[0]
    out = np.zeros(rows * 32)
    for r in range(rows):
        for j in range(32):
            off = r * 128 + (((j >> 2) ^ (r & 7)) << 4) + 4 * (j & 3)
            out[off // 4] = raw[r, j]
    return out


def test_gin2_rotated_swizzled_reads_build_the_band_rows_conflict_free():
    rng = np.random.default_rng(7)
    raw = rng.standard_normal((RAW_ROWS, KKC))
    smem = swizzle128_store(raw)
    for q in range(4):
        banks = {}                                        # (c0, t) -> banks touched by the 32 lanes of the warp
        for lane in range(32):
            xl, h = 32 * q + lane, lane >> 3
            base_n = [(xl + 31 - ((n + h) & 3)) * 128 + 4 * ((n + h) & 3) for n in range(4)]
            e_n = [((xl + 31 - ((n + h) & 3)) & 7) << 4 for n in range(4)]
            rot1, rot2 = bool(h & 1), bool(h & 2)
            row = np.zeros(32)
            for c0 in (0, 16):
                w = [0.0] * 16
                for t in range(16):
                    jj = c0 + t
                    mm, n = jj >> 2, jj & 3
                    km = (mm ^ (4 * (mm & 1))) << 4
                    off = base_n[n] - 512 * mm + (km ^ e_n[n])
                    assert off % 4 == 0 and 0 <= off // 4 < smem.size
                    banks.setdefault((c0, t), []).append((off // 4) % 32)
                    w[t] = smem[off // 4]
                for gq in range(0, 16, 4):                # rotate every group of 4 back by h (two select stages)
                    a0, a1 = (w[gq + 3], w[gq + 0]) if rot1 else (w[gq + 0], w[gq + 1])
                    a2, a3 = (w[gq + 1], w[gq + 2]) if rot1 else (w[gq + 2], w[gq + 3])
                    w[gq + 0], w[gq + 1] = (a2, a3) if rot2 else (a0, a1)
                    w[gq + 2], w[gq + 3] = (a0, a1) if rot2 else (a2, a3)
                row[c0:c0 + 16] = w
            want = np.array([raw[xl + 31 - jj, jj] for jj in range(32)])   # Gd[xl][jj] = raw[xl + 31 - jj][jj]
            assert np.array_equal(row, want), (q, lane)
        for key, b in banks.items():
            assert len(set(b)) == 32, f"bank conflict in warp quarter {q}, instruction {key}"


@pytest.mark.parametrize("P", [17, 40, 100, 192, 193])
@pytest.mark.parametrize("groups", [2, 3])
def test_box_release_closed_form_matches_the_search_loop(P, groups):
    """A builder warp visits chunks k, k + groups, ... of a tile; box b of the g slice is last read by chunk
    min(NKC-1, b + koff).  The closed form used by the kernel must release exactly what the original loop released."""
    m = modes(P)
    nkc, nb, koff = m["NKC"], m["n_gboxes"], m["koff"]
    for first in range(groups):
        next_a = next_b = 0
        for k in range(first, nkc, groups):
            last_visit = k + groups >= nkc
            while next_a < nb:                            # round-1 formulation
                kl = min(next_a + koff, nkc - 1)
                if not (kl <= k + groups - 1 or last_visit):
                    break
                next_a += 1
            target = nb if last_visit else min(nb, k + groups - koff)   # round-2 closed form
            next_b = max(next_b, target)
            assert next_a == next_b, (P, groups, first, k)
            # nothing this warp still reads may be released: chunk kk needs the planes 32 kk - delta - 127 .. 32 kk + 31 - delta
            delta = m["delta"][0]
            for kk in range(k + groups, nkc, groups):
                lowest_needed = max(0, (KKC * kk - delta - (KTM - 1)) // 32)
                assert next_b <= lowest_needed, (P, groups, first, k, kk)
        assert next_b == nb                               # every box handed back once per tile


def test_split_teams_never_lap_the_band_ring():
    """Two teams take the chunks alternately; a team that started on the ring's SECOND phase (team index >= ring depth)
    would see the fresh barrier's phase-1 test pass at once -- the configuration check keeps band_slots >= teams."""
    teams = 2
    for band_slots in range(2, 9):
        for team in range(teams):
            assert team // band_slots == 0              # first wait of every team is on phase 0 of its slot
