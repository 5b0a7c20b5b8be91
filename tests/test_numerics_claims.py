"""CPU checks of two arithmetic identities the CUDA kernels rely on.  No GPU, no oracle."""
import numpy as np


def test_packed_fp16_bias_add_equals_the_reference_rounding_path():
    """GuidanceNet epilogues add the bias as ONE fp16 add (HADD2).  The reference path (ATen/cuDNN fp16 conv) computes
    half(float(half(acc)) + float(bias)): an fp32 add of two fp16 values, rounded to fp16.  fp32 carries 24 >= 2*11 + 2
    significant bits, so that double rounding is innocuous and equals the correctly rounded fp16 sum — checked here against an
    exact (float64) sum rounded once, over every exponent pairing, random mantissas and the tie cases."""
    rs = np.random.default_rng(7)
    bits = rs.integers(0, 1 << 16, size=(1 << 20, 2), dtype=np.uint32).astype(np.uint16)
    a, b = bits[:, 0].view(np.float16), bits[:, 1].view(np.float16)
    ok = np.isfinite(a) & np.isfinite(b)
    a, b = a[ok], b[ok]
    # ties: a + b exactly half way between two fp16 values (b = half an ulp of a)
    m = rs.integers(0, 1 << 10, size=4096).astype(np.uint16)
    for e in range(2, 30):
        big = ((np.uint16(e) << 10) | m).view(np.float16)
        half_ulp = np.float16(2.0 ** (e - 15 - 10 - 1))
        a = np.concatenate([a, big, big])
        b = np.concatenate([b, np.full(big.shape, half_ulp, np.float16), np.full(big.shape, -half_ulp, np.float16)])
    with np.errstate(over="ignore"):
        via_f32 = (a.astype(np.float32) + b.astype(np.float32)).astype(np.float16)      # the reference's path
        exact_once = (a.astype(np.float64) + b.astype(np.float64)).astype(np.float16)   # correctly rounded fp16 add
    assert a.size > 900000
    assert np.array_equal(via_f32.view(np.uint16), exact_once.view(np.uint16))


def test_magic_add_coordinates_are_the_floor_at_every_width():
    """Fused-index marcher: asuint(fadd.rd(p, 2^(23-n))) = exponent bits | floor(p * 2^n) for p in [0, 1), n <= 11 — the host
    statement of coord_bits_at (csrc/rto_ray.cuh).  Emulated with float64 (exact sum, explicit floor to the fp32 grid)."""
    rs = np.random.default_rng(11)
    p = np.concatenate([rs.random(200000, dtype=np.float32), np.float32([0.0, 1.0 - 1e-6, 0.5, 2.0 ** -23, 1.0 - 2.0 ** -24])])
    p = np.minimum(p, np.float32(1.0 - 1e-6))
    for n in range(1, 12):
        M = 2.0 ** (23 - n)
        s = p.astype(np.float64) + M                       # exact in float64
        rd = np.floor(s * 2.0 ** n) / 2.0 ** n             # round toward -inf to the fp32 grid of [M, 2M): spacing 2^-n
        got = rd.astype(np.float32).view(np.uint32)
        want = (np.uint32((127 + 23 - n) << 23) | np.floor(p.astype(np.float64) * 2.0 ** n).astype(np.uint32))
        assert np.array_equal(got, want), n
        # the level-n coordinate is the top n bits of the 23-bit coordinate the tree walker uses
        c23 = np.floor(p.astype(np.float64) * 2.0 ** 23).astype(np.uint32)
        assert np.array_equal(want & np.uint32((1 << n) - 1), c23 >> np.uint32(23 - n))
