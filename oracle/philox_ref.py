"""numpy restatement of the DEVICE random stream ("JNE2")  --  TEST INFRASTRUCTURE ONLY.

Philox4x32-10 (Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3",
SC'11; Random123 reference constants) keys one xoshiro128++ substream (Blackman, Vigna,
"Scrambled linear pseudorandom number generators", 2021) per (row, epoch of 128 steps, half),
whose words feed the Box-Muller map of johansen_null_eigenspectra_b200/csrc/jne_rng.cuh.  The
uniform words are bit-exact with the device; the normals agree to MUFU approximation error
(~1e-6), so tests compare them with a stated tolerance.  This replaces reference function gen_normal_matrix
(src/rng_matrix.rs:11-37), whose Xoshiro256++/ziggurat stream is machine-dependent and is
not reproduced.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
KEY1 = 0x4A4E4532  # "JNE2": second key word, fixed
EPOCH_BLOCKS = 32   # four-step blocks per epoch (128 steps); every other block belongs to the same substream


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


def _rotl(x, r):
    return ((x << np.uint32(r)) | (x >> np.uint32(32 - r))).astype(np.uint32)


def xoshiro128pp(state):
    """One step of xoshiro128++ on a list of four uint32 arrays (updated in place); returns the output words."""
    s0, s1, s2, s3 = state
    with np.errstate(over="ignore"):
        out = (_rotl((s0 + s3).astype(np.uint32), 7) + s0).astype(np.uint32)
    t = (s1 << np.uint32(9)).astype(np.uint32)
    s2 = s2 ^ s0
    s3 = s3 ^ s1
    s1 = s1 ^ s2
    s0 = s0 ^ s3
    s2 = s2 ^ t
    s3 = _rotl(s3, 11)
    state[:] = [s0, s1, s2, s3]
    return out


def stream_words(dim, steps, seed):
    """The uniform words of the stream as a (dim, nblocks, 4) uint32 array: block b = t >> 2 of row r is block
    j = (b & 31) >> 1 of the substream (r, epoch e = b >> 5, half h = b & 1), whose generator state is
    Philox4x32-10(key = (seed, KEY1), ctr = (e, r, h, 0)) and whose outputs 4j .. 4j+3 are the block's words."""
    nb = (steps + 3) // 4
    ne = (nb + EPOCH_BLOCKS - 1) // EPOCH_BLOCKS
    e = np.arange(ne, dtype=np.uint32)[None, :, None]
    r = np.arange(dim, dtype=np.uint32)[:, None, None]
    h = np.arange(2, dtype=np.uint32)[None, None, :]
    state = [w.copy() for w in philox4x32_10(e, r, h, 0, np.uint32(seed), np.uint32(KEY1))]
    zero = (state[0] | state[1] | state[2] | state[3]) == 0
    state[0][zero] = 1                                  # the generator's one forbidden state
    words = np.empty((dim, ne, EPOCH_BLOCKS // 2, 2, 4), dtype=np.uint32)   # [row, epoch, j, half, word]
    for j in range(EPOCH_BLOCKS // 2):
        for w in range(4):
            words[:, :, j, :, w] = xoshiro128pp(state)
    return words.reshape(dim, ne * EPOCH_BLOCKS, 4)[:, :nb]


def normal_matrix(dim, steps, seed):
    """d x T matrix of float32 normals (returned as float64): words (0,1) of a block -> steps 4b, 4b+1 (cos, sin),
    words (2,3) -> steps 4b+2, 4b+3."""
    w = stream_words(dim, steps, seed)
    za, zb = box_muller(w[..., 0], w[..., 1])
    zc, zd = box_muller(w[..., 2], w[..., 3])
    z = np.stack([za, zb, zc, zd], axis=-1).reshape(dim, -1)[:, :steps]
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
    st = [np.array([v], dtype=np.uint32) for v in (1, 2, 3, 4)]   # rand_xoshiro's xoshiro128++ reference vector
    print([int(xoshiro128pp(st)[0]) for _ in range(4)], "expected [641, 1573767, 3222811527, 3517856514]")
