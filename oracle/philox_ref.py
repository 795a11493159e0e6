"""numpy restatement of the DEVICE random stream  --  TEST INFRASTRUCTURE ONLY.

Philox4x32-10 (Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3",
SC'11; Random123 reference constants) and the Box-Muller map used by
johansen_null_eigenspectra_b200/csrc/jne_rng.cuh.  The uniform words are bit-exact with the
device; the normals agree to MUFU approximation error (~1e-6), so tests compare them with
a stated tolerance.  This replaces reference function gen_normal_matrix
(src/rng_matrix.rs:11-37), whose Xoshiro256++/ziggurat stream is machine-dependent and is
not reproduced.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
KEY1 = 0x4A4E4531  # "JNE1": second key word, fixed


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32 arrays."""
    c0, c1, c2, c3, k0, k1 = [np.asarray(x, dtype=np.uint32) for x in (c0, c1, c2, c3, k0, k1)]
    c0, c1, c2, c3, k0, k1 = np.broadcast_arrays(c0, c1, c2, c3, k0, k1)
    c0, c1, c2, c3, k0, k1 = [x.copy() for x in (c0, c1, c2, c3, k0, k1)]
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for r in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & mask).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & mask).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            if r < 9:
                k0 = (k0 + W0).astype(np.uint32)
                k1 = (k1 + W1).astype(np.uint32)
    return c0, c1, c2, c3


def box_muller(wa, wb):
    """(wa, wb) uint32 -> two float32 normals, as jne_rng.cuh::box_muller:
    u = (wa + 0.5) 2^-32 in (0,1], r = sqrt(-2 ln u); theta = 2 pi * int32(wb) 2^-32 in [-pi, pi)."""
    wa = np.asarray(wa, dtype=np.uint32)
    wb = np.asarray(wb, dtype=np.uint32)
    u = np.float32(wa.astype(np.float32)) * np.float32(2.0 ** -32) + np.float32(2.0 ** -33)
    u = u.astype(np.float32)
    # the device multiplies lg2(u) by -0x1.62e436p+0 (= -2 ln 2, three float ulps further out: its variance
    # calibration, see jne_rng.cuh); this mirror follows it
    r = np.sqrt(-float.fromhex("0x1.62e436p+0") * np.log2(u.astype(np.float64))).astype(np.float32)
    th = wb.view(np.int32).astype(np.float32).astype(np.float64) * (2.0 ** -32) * 2.0 * np.pi
    return (r * np.cos(th)).astype(np.float32), (r * np.sin(th)).astype(np.float32)


def normal_matrix(dim, steps, seed):
    """d x T matrix of float32 normals (returned as float64), element (r, t) drawn from
    Philox(key=(seed, KEY1), ctr=(t >> 2, r, 0, 0)): words (0,1) -> steps 4b, 4b+1 (cos, sin),
    words (2,3) -> steps 4b+2, 4b+3."""
    nb = (steps + 3) // 4
    tb = np.arange(nb, dtype=np.uint32)[None, :]
    rows = np.arange(dim, dtype=np.uint32)[:, None]
    w0, w1, w2, w3 = philox4x32_10(tb, rows, 0, 0, np.uint32(seed), np.uint32(KEY1))
    za, zb = box_muller(w0, w1)
    zc, zd = box_muller(w2, w3)
    z = np.stack([za, zb, zc, zd], axis=-1).reshape(dim, nb * 4)[:, :steps]
    return z.astype(np.float64)


if __name__ == "__main__":
    # Random123 known-answer vectors (kat_vectors, philox4x32 10 rounds)
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, exp in kat:
        out = tuple(int(x) for x in philox4x32_10(*ctr, *key))
        print([hex(x) for x in out], "OK" if out == exp else "MISMATCH")
