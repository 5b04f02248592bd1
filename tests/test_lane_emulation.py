"""Host-side emulation of the lane-level algorithms of two kernels (no GPU): the index algebra the CUDA code
implements, checked in numpy against the plain definitions.  These emulations were written BEFORE the kernels and
are what their first GPU runs were debugged against (profiles/r02_summary.md); the formulas below are the ones in
fastoverlap_b200/csrc/fo_spherical.cu (d2_sw / d2_pos / d2_rowperm, sph_direct2_kernel's fragment addressing and
chained second product; FftLane::run and the step A / step B packing of sph_isoft5_kernel).

DMMA.8x8x4 fragment layout (g = lane >> 2, t = lane & 3): A[row g][k t], B[k t][col g], C[row g][cols 2t, 2t + 1]."""
import numpy as np
import pytest


# ----------------------------------------------------------------------------- sph_direct2_kernel

def d2_sw(nt, p):
    m = nt & 3
    return (p & 3) if m == 0 else (((p >> 1) & 1) if m == 2 else 0)


def d2_pos(nt, p, c):
    return p * 8 * nt + (((c >> 3) ^ d2_sw(nt, p)) << 3) + (c & 7)


def d2_rowperm(k):
    return (k & ~7) | ((k & 1) << 2) | ((k & 7) >> 1)


def dmma(c, a, b):
    """c[32][2] += the m8n8k4 product of the per-lane operands a[32], b[32]."""
    A = np.zeros((8, 4))
    B = np.zeros((4, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        A[g, t] = a[lane]
        B[t, g] = b[lane]
    D = A @ B
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        c[lane, 0] += D[g, 2 * t]
        c[lane, 1] += D[g, 2 * t + 1]


@pytest.mark.parametrize("nt", range(1, 17))
def test_swizzled_fragment_loads_are_conflict_free(nt):
    """The four k rows x eight columns of a fragment load hit 32 distinct banks for every tile count (pitch 8 nt
    doubles, no padding columns), and d2_pos is a bijection of the block."""
    lanes = np.arange(32)
    g, t4 = lanes >> 2, lanes & 3
    sw = np.array([d2_sw(nt, t) for t in t4])
    for rt in range(nt):
        for base in (0, 4, 8, 12):
            addr = (base + t4) * 8 * nt + ((rt ^ sw) << 3) + g
            assert len(set(addr % 32)) == 32
    pos = {d2_pos(nt, p, c) for p in range(16) for c in range(8 * nt)}
    assert pos == set(range(16 * 8 * nt))
    assert sorted(d2_rowperm(k) for k in range(64)) == list(range(64))


@pytest.mark.parametrize("natoms,L", [(38, 15), (13, 7), (8, 4), (21, 9), (3, 2)])
def test_direct2_layouts_and_chained_products(natoms, L):
    """Producers' operand images + the consumer's two DMMA products (T's C fragments reused as the A fragments of
    the second product over the k-steps {2 t + h}) + the epilogue give I[l, m1, m2] = sum_jk Y^A_lm1(j) B_l[j, k]
    conj Y^B_lm2(k) for all |m1| <= l, 0 <= m2 <= l."""
    rng = np.random.default_rng(natoms * 100 + L)
    N8 = (natoms + 7) & ~7
    NCT = N8 >> 3
    lanes = np.arange(32)
    g, t4 = lanes >> 2, lanes & 3
    for l in range(L + 1):
        nrt = (2 * (l + 1) + 7) >> 3
        R8 = 8 * nrt
        YA = rng.normal(size=(natoms, l + 1)) + 1j * rng.normal(size=(natoms, l + 1))
        YB = rng.normal(size=(natoms, l + 1)) + 1j * rng.normal(size=(natoms, l + 1))
        Bl = rng.normal(size=(natoms, natoms))
        A1 = np.full(N8 * R8, np.nan)
        B2 = np.full(N8 * R8, np.nan)
        B1 = np.full(N8 * N8, np.nan)
        for atom in range(N8):           # sph_prep2_kernel<false / true>: thread (row, complex column m)
            for m in range(4 * nrt):
                va = YA[atom, m] if (atom < natoms and m <= l) else 0
                vb = YB[atom, m] if (atom < natoms and m <= l) else 0
                pa = d2_pos(nrt, atom, 2 * m)
                A1[pa], A1[pa + 1] = np.real(va), np.imag(va)
                pb = d2_pos(nrt, d2_rowperm(atom), 2 * m)
                B2[pb], B2[pb + 1] = np.real(vb), np.imag(vb)
        for j in range(N8):              # sph_bessel2_kernel
            for k in range(N8):
                B1[d2_pos(NCT, j, k)] = Bl[j, k] if (j < natoms and k < natoms) else 0
        assert not (np.isnan(A1).any() or np.isnan(B2).any() or np.isnan(B1).any())
        T = YA.T @ Bl
        ref = np.zeros((2 * l + 1, l + 1), complex)
        for m1 in range(-l, l + 1):
            Tm = T[m1] if m1 >= 0 else (-1) ** (-m1) * np.conj(T[-m1])
            ref[m1 + l] = [np.sum(Tm * np.conj(YB[:, m2])) for m2 in range(l + 1)]
        out = np.full((2 * l + 1, l + 1), np.nan, complex)
        swB = np.array([d2_sw(NCT, t) for t in t4])
        swY = np.array([d2_sw(nrt, t) for t in t4])
        for rt1 in range(nrt):
            c1 = np.zeros((NCT, 32, 2))
            for ks in range(2 * NCT):
                av = A1[(t4 + 4 * ks) * R8 + ((rt1 ^ swY) << 3) + g]
                for ct in range(NCT):
                    dmma(c1[ct], av, B1[(t4 + 4 * ks) * N8 + ((ct ^ swB) << 3) + g])
            for rt2 in range(nrt):
                c2 = np.zeros((32, 2))
                for ct in range(NCT):
                    for h in range(2):
                        dmma(c2, c1[ct][:, h], B2[(8 * ct + 4 * h + t4) * R8 + ((rt2 ^ swY) << 3) + g])
                for lane in range(32):
                    gg, tt = lane >> 2, lane & 3
                    px0, px1 = c2[lane ^ 4]
                    m1, m2 = (rt1 * 8 + gg) >> 1, rt2 * 4 + tt
                    if m1 > l or m2 > l:
                        continue
                    if gg & 1 == 0:
                        out[l + m1, m2] = complex(c2[lane, 0] + px1, px0 - c2[lane, 1])
                    elif m1 > 0:
                        sg = -1.0 if m1 & 1 else 1.0
                        out[l - m1, m2] = complex(sg * (px0 - c2[lane, 1]), sg * (-c2[lane, 0] - px1))
        assert not np.isnan(out).any()
        assert np.abs(out - ref).max() < 1e-11 * max(1.0, np.abs(ref).max())


def test_packed_coefficient_index_of_the_epilogue():
    """sph_direct2_kernel's packed index t = s^2 + (m2 == s ? a : s + 1 + m2), s = max(a, m2), is the inverse of
    sph_ipack_kernel's unpacking of the shell-ordered entries."""
    for t in range(16 * 16):
        s = int(np.sqrt(t))
        while (s + 1) * (s + 1) <= t:
            s += 1
        while s * s > t:
            s -= 1
        q = t - s * s
        a, m2 = (q, s) if q <= s else (s, q - s - 1)
        sh = max(a, m2)
        assert sh * sh + (a if m2 == sh else sh + 1 + m2) == t


# ----------------------------------------------------------------------------- sph_isoft5_kernel

def fft4(x0, x1, x2, x3):
    a, b, c, d = x0 + x2, x0 - x2, x1 + x3, 1j * (x1 - x3)
    return [a + c, b + d, a - c, b - d]


def fft32_four_lanes(x):
    """FftLane::run for the four lanes of a quad: lane n1 = 2 b1 + b0 brings the inputs 4 n2 + n1."""
    r2 = 1 / np.sqrt(2)
    w, w3 = (1 + 1j) * r2, (-1 + 1j) * r2
    Z = {}
    for n1 in range(4):
        b0, b1 = n1 & 1, (n1 >> 1) & 1
        xi = [x[4 * n2 + n1] for n2 in range(8)]
        for i in range(8):               # sign flips of the inputs (xor of the sign bits in the kernel)
            s = (-1 if (b1 and (i & 1)) else 1) * (-1 if (b0 and (i & 2)) else 1)
            xi[i] = s * xi[i]
        E = fft4(xi[0], xi[2], xi[4], xi[6])
        O = fft4(xi[1], xi[3], xi[5], xi[7])
        M = [1, w, 1j, w3] if not b0 else [1j, w3, 1, w]
        z = [None] * 8
        for k in range(4):
            z[k], z[k + 4] = E[k] + M[k] * O[k], E[k] - M[k] * O[k]
        r = 2 * b0 + 4 * b1              # z[k] = y[k ^ r]: every lane keeps / sends the same registers
        Z[n1] = [z[k] * np.exp(2j * np.pi * n1 * (k ^ r) / 32) for k in range(8)]
    out = np.zeros(32, complex)
    for n1 in range(4):
        b0, b1 = n1 & 1, (n1 >> 1) & 1
        for jj in range(2):
            F = fft4(Z[n1][jj], Z[n1 ^ 1][jj + 2], Z[n1 ^ 2][jj + 4], Z[n1 ^ 3][jj + 6])
            for k1 in range(4):
                idx, fac = k1, 1
                if b0:
                    idx, fac = (-k1) & 3, 1j ** k1
                if b1:
                    fac = fac * (-1) ** k1
                out[2 * n1 + jj + 8 * k1] = fac * F[idx]
    return out


def test_four_lane_fft32():
    rng = np.random.default_rng(32)
    x = rng.normal(size=32) + 1j * rng.normal(size=32)
    ref = np.array([np.sum(x * np.exp(2j * np.pi * np.arange(32) * k / 32)) for k in range(32)])
    assert np.abs(fft32_four_lanes(x) - ref).max() < 1e-12


def test_isoft5_plane_transform():
    """Block [m1 mod 32][m2 = 0..15] -> step A (columns, row 16 = 0) -> step B (rows al, al + 16 packed into one
    complex transform over m2 = -15..15, V(-m2) = conj V(m2)) = the real grid sum_{m1, m2} S e^{i (m1 alpha + m2 gamma)}."""
    rng = np.random.default_rng(5)
    L, F = 15, 32
    S = rng.normal(size=(2 * L + 1, L + 1)) + 1j * rng.normal(size=(2 * L + 1, L + 1))
    for a in range(1, L + 1):            # the m2 = 0 column of a real grid is Hermitian in m1
        S[L - a, 0] = np.conj(S[L + a, 0])
    S[L, 0] = S[L, 0].real
    ang = np.arange(F) * 2 * np.pi / F
    ref = np.zeros((F, F))
    for i1, m1 in enumerate(range(-L, L + 1)):
        for m2 in range(L + 1):
            term = S[i1, m2] * np.exp(1j * (m1 * ang[:, None] + m2 * ang[None, :]))
            ref += term.real if m2 == 0 else 2 * term.real
    blk = np.zeros((32, 16), complex)
    for i1, m1 in enumerate(range(-L, L + 1)):
        blk[m1 % 32] = S[i1]
    V = np.zeros((32, 16), complex)
    for m2 in range(16):
        col = blk[:, m2].copy()
        col[16] = 0
        V[:, m2] = fft32_four_lanes(col)
    g = np.zeros((F, F))
    rows = []
    for f in range(16):
        a0 = (f & ~3) | ((f & 1) << 1) | ((f >> 1) & 1)   # neighbouring transforms two rows apart (banks)
        rows.append(a0)
        Z = np.zeros(32, complex)
        for n in range(32):
            if n <= 15:
                Z[n] = V[a0, n] + 1j * V[a0 + 16, n]
            elif n > 16:
                Z[n] = np.conj(V[a0, 32 - n]) + 1j * np.conj(V[a0 + 16, 32 - n])
        o = fft32_four_lanes(Z)
        g[a0], g[a0 + 16] = o.real, o.imag
    assert sorted(rows) == list(range(16))
    assert np.abs(g - ref).max() < 1e-11 * np.abs(ref).max()


def test_k5_groups_dealt_evenly_to_subpartitions():
    """The nibble table of sph_isoft5_kernel is a permutation of the 16 entry groups (32 lanes = 16 shell-ordered
    entries x level parity) and evens out the level trips of the four warps of every SM sub-partition (warp & 3)."""
    import math
    L = 15

    def trips(t, par):
        s = math.isqrt(t)
        l0 = s + ((s ^ par) & 1)          # first level >= shell of the lane's parity
        return 0 if l0 > L else (L - l0) // 2 + 1

    c = [max(trips(t, p) for t in range(16 * w, 16 * w + 16) for p in (0, 1)) for w in range(16)]
    assert c == [8, 6, 6, 5, 4, 4, 4, 3, 3, 2, 2, 2, 2, 1, 1, 1]
    tab = [(0xfdce98ba65473210 >> (4 * w)) & 15 for w in range(16)]
    assert sorted(tab) == list(range(16))
    assert [sum(c[tab[w]] for w in range(s, 16, 4)) for s in range(4)] == [14, 14, 14, 12]
    assert [sum(c[w] for w in range(s, 16, 4)) for s in range(4)] == [17, 13, 13, 11]
