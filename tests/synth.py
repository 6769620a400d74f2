"""Deterministic test inputs that do not depend on numpy's RNG streams (pure integer math)."""
import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_G = np.uint64(0x9E3779B97F4A7C15)


def _mix(z):
    z = (z ^ (z >> np.uint64(30))) * _M1
    z = (z ^ (z >> np.uint64(27))) * _M2
    return z ^ (z >> np.uint64(31))


def synth_bytes(seed, n, kind="rand"):
    """n bytes: rand | nib (4-bit entropy) | bit | zero | text | p<k> (period k of random bytes) | rec (LZ-compressible records)"""
    if kind == "zero":
        return np.zeros(n, dtype=np.uint8)
    if kind == "rec":
        return synth_records(seed, n)
    if kind.startswith("p") and kind[1:].isdigit():
        k = int(kind[1:])
        unit = synth_bytes(seed, k, "rand")
        return np.tile(unit, n // k + 1)[:n].copy()
    words = (n + 7) // 8
    with np.errstate(over="ignore"):
        idx = np.arange(1, words + 1, dtype=np.uint64)
        z = _mix(np.uint64(seed) * _G + idx * _G)
    b = z.view(np.uint8)[:n].copy()
    if kind == "rand":
        return b
    if kind == "nib":
        return b & np.uint8(0x0F)
    if kind == "bit":
        return b & np.uint8(1)
    if kind == "text":
        alphabet = np.frombuffer(b"eeeeeeee tttttt aaaaa ooooo iiii nnnn ssss hhh rrr dd ll cu\nmwfgyp", dtype=np.uint8)
        return alphabet[b & np.uint8(63)]
    raise ValueError(kind)


def synth_records(seed, n):
    """LZ-compressible bytes: phrases drawn from a 4 KiB text vocabulary, sprinkled with random bytes, long byte runs and
    (every ~3 MiB) a 300 KiB copy of data from more than 2 MiB back — exercises matches, repcodes, long lengths, RLE blocks
    and the window limit of the block codecs"""
    vocab = synth_bytes(seed + 1, 4096, "text").tobytes()
    noise = synth_bytes(seed + 2, 65536, "rand").tobytes()
    out = bytearray()
    i = 0
    next_far = 3 << 20
    mask = (1 << 64) - 1
    while len(out) < n:
        z = (seed * 0x9E3779B97F4A7C15 + (i + 1) * 0xD1B54A32D192ED03) & mask
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & mask
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & mask
        z ^= z >> 31
        a = z % 4000
        out += vocab[a:a + 4 + ((z >> 12) % 60)]
        if (z >> 20) % 4 == 0:
            b = (z >> 24) % 65000
            out += noise[b:b + (z >> 44) % 24]
        if (z >> 30) % 512 == 0:
            out += bytes([(z >> 40) & 255]) * ((z >> 48) % 200000)
        if len(out) >= next_far:
            src = len(out) - (2 << 20) - 400000
            out += out[src:src + 300000]
            next_far += 3 << 20
        i += 1
    return np.frombuffer(bytes(out[:n]), dtype=np.uint8).copy()


def chunker_params(target):
    """src/longtail.c:1985-1987 with GetMinChunkSize() == 48"""
    return max(48, target // 8), max(48, target // 2), max(48, target * 2)


def small_tree(target=256):
    """the asset tree of SURVEY.md F12: sizes around every edge of the part / window rules"""
    part = target * 1024
    dup = synth_bytes(11, 300000)
    assets = [
        ("a/empty.bin", synth_bytes(1, 0)), ("a/one.bin", synth_bytes(2, 1)), ("a/47.bin", synth_bytes(3, 47)),
        ("a/48.bin", synth_bytes(4, 48)), ("a/49.bin", synth_bytes(5, 49)), ("a/300.bin", synth_bytes(6, 300)),
        ("b/exact.bin", synth_bytes(7, part)), ("b/two.bin", synth_bytes(8, part + 12345)),
        ("b/four.bin", synth_bytes(9, 3 * part + 777)), ("c/dup1.bin", dup), ("d/dup2.bin", dup.copy()),
        ("e/low.bin", synth_bytes(12, 500000, "nib")), ("emptydir/", synth_bytes(13, 0)),
        ("t/text.bin", synth_bytes(14, 400000, "text")), ("z/zero.bin", synth_bytes(15, 200000, "zero")),
    ]
    assets.sort(key=lambda a: a[0].encode())
    return assets
