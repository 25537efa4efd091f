"""Host-side check of the 16-bit fixed-point code the predictor kernels use for the saved SiLU derivatives
(gaudi_b200/csrc/tc_common.cuh: SV_FX -- n = round(52428 d + 6554), built as the low mantissa bits of fma(d, 52428, 2^23 + 6554)
and decoded as (y - (2^23 + 6554)) / 52428).  numpy restatement of the same fp32 arithmetic: the magic-number construction is
round-to-nearest, 0 is exact, the error bound quoted in DESIGN.md section 3 holds on the whole range of SiLU'."""
import numpy as np

SCALE = np.float32(52428.0)
BIAS = np.float32(8388608.0 + 6554.0)
INV = np.float32(1.0) / SCALE


def encode(d):
    y = (d.astype(np.float64) * np.float64(SCALE) + np.float64(BIAS)).astype(np.float32)      # one fp32 FMA: exact product, one rounding
    return (y.view(np.uint32) & np.uint32(0xFFFF)).astype(np.uint16)


def decode(n):
    y = (np.uint32(0x4B000000) | n.astype(np.uint32)).view(np.float32)                           # PRMT: [code | exponent of 2^23]
    return (y - BIAS) * INV


def silu_prime(x):
    s = 1.0 / (1.0 + np.exp(-x))
    return s * (1.0 + x * (1.0 - s))


def test_code_is_round_to_nearest_and_zero_is_exact():
    d = np.linspace(-0.0998, 1.0998, 200001, dtype=np.float32)
    n = encode(d)
    assert np.array_equal(n, np.rint(d.astype(np.float64) * 52428.0 + 6554.0).astype(np.uint16))
    assert int(n.min()) > 0 and int(n.max()) < 65535                                            # the whole range fits 16 bits
    assert decode(encode(np.zeros(4, np.float32))).tolist() == [0.0, 0.0, 0.0, 0.0]
    assert abs(float(decode(encode(np.ones(1, np.float32)))[0]) - 1.0) <= 1.2e-7


def test_error_bound_on_the_range_of_the_silu_derivative():
    x = np.linspace(-30.0, 30.0, 600001)
    d = silu_prime(x).astype(np.float32)
    assert -0.1 < float(d.min()) and float(d.max()) < 1.1
    err = np.abs(decode(encode(d)).astype(np.float64) - d.astype(np.float64))
    assert float(err.max()) <= 9.6e-6                                                           # half a step of 1 / 52428
    h = np.abs(d.astype(np.float16).astype(np.float64) - d.astype(np.float64))                  # what fp16 storage would do
    assert float(h.max()) >= 2e-4
