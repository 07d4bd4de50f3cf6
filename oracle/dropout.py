"""Keep masks of the CUDA path's counter-based dropout, restated on the host (numpy integer arithmetic).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

This is NOT reference behaviour: the reference (dpr_scale/models/hf_model.py:22-24 -> HF `nn.Dropout`,
site-packages/transformers/models/bert/modeling_bert.py:111, :296, :354 and the attention-probability dropout) draws
torch's Philox masks, which no other implementation reproduces.  What IS pinned to the reference is the arithmetic
around the mask (tests/test_dropout_gpu.py replays OUR masks inside the HF oracle).  This module pins the definition
of our masks themselves - the one include/dprb.h states and dpr_scale_b200/csrc/common.cuh (`Drop`, `make_drop`)
implements - bit for bit, so the mask generator is integer work with an independent checker like the rest of the path.

  element (r, c) of a site is KEPT iff lane16 >= thresh16, where
    site_seed = fold32(seed + (layer * 8 + site + 1) * 0x9E3779B97F4A7C15)
    x   = (r * row_mul) * 0x9E3779B1 + (c / 8) * 0x85EBCA6B + site_seed         (mod 2^32)   one chain per 8 columns
    x  ^= x >> 16;  x *= 0x7FEB352D;  x ^= x >> 15
    h   = x * K[(c / 2) % 4];  h ^= h >> 16                                    (mod 2^32)   one finaliser per pair
    lane16 = low half of h for even c, high half for odd c;  thresh16 = round(p * 65536), kept values scale by
    1 / (1 - thresh16 / 65536).
"""
import numpy as np

M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF
PAIR_MULTIPLIERS = (0x846CA68B, 0xC2B2AE35, 0x27D4EB2F, 0x165667B1)


def site_seed32(seed: int, layer: int, site: int) -> int:
    s64 = (int(seed) + (layer * 8 + site + 1) * 0x9E3779B97F4A7C15) & M64
    return (s64 ^ (s64 >> 32)) & M32


def thresh16(p: float) -> int:
    p32 = float(np.float32(p))                       # the C ABI takes a float
    t = int(p32 * 65536.0 + 0.5) if p32 > 0 else 0
    return min(t, 65535)


def scale(p: float) -> float:
    t = thresh16(p)
    return float(np.float32(1.0) / (np.float32(1.0) - np.float32(t) / np.float32(65536.0))) if t else 1.0


def keep_mask(rows: int, cols: int, p: float, seed: int, layer: int, site: int, row_mul: int = 1) -> np.ndarray:
    """uint8 [rows, cols]: 1 = kept.  Same arguments as ``dprb_dropout_mask`` (row_mul: the pruned last layer keys its
    CLS rows by row * S)."""
    t = thresh16(p)
    if t == 0:
        return np.ones((rows, cols), dtype=np.uint8)
    sd = np.uint64(site_seed32(seed, layer, site))
    r = np.arange(rows, dtype=np.uint64)[:, None]
    c = np.arange(cols, dtype=np.uint64)[None, :]
    m32 = np.uint64(M32)
    x = (((r * np.uint64(row_mul)) & m32) * np.uint64(0x9E3779B1) + (c >> np.uint64(3)) * np.uint64(0x85EBCA6B) + sd) & m32
    x = x ^ (x >> np.uint64(16))
    x = (x * np.uint64(0x7FEB352D)) & m32
    x = x ^ (x >> np.uint64(15))
    k = np.array(PAIR_MULTIPLIERS, dtype=np.uint64)[((c >> np.uint64(1)) & np.uint64(3)).astype(np.int64)]
    h = (x * k) & m32
    h = h ^ (h >> np.uint64(16))
    lane = np.where((c & np.uint64(1)) == 1, h >> np.uint64(16), h & np.uint64(0xFFFF))
    return (lane >= np.uint64(t)).astype(np.uint8)
