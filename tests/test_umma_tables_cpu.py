"""CPU: the host-built tables of the tensor-core FIR (f9_umma.cu).  f9_umma_selfcheck plans a ratio exactly as the
library does at run time and verifies the fp16 head/tail weight tiles against the fp32 polyphase weights tap by tap
(every tap exactly once, zeros elsewhere), so a table defect is caught here without a GPU."""
import ctypes as C

import pytest

RATIOS = [(320, 147), (147, 160), (1, 4), (2, 1), (4, 1), (147, 80), (147, 320), (160, 147), (1, 1), (3, 2), (2, 3), (441, 160)]


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
@pytest.mark.parametrize("pq", RATIOS)
def test_tables_reproduce_polyphase_weights(f9, kind, pq):
    info = (C.c_int * 8)()
    err = f9.lib().f9_umma_selfcheck(kind, pq[0], pq[1], info)
    if err == -2.0:
        pytest.skip("no tensor-core plan for this ratio: served by the CUDA-core kernels")
    assert err >= 0.0, f"table defect {err} for kind {kind} ratio {pq}"
    # x = x0 + x1/2048 with fp16 parts: |w - (w0 + w1/2048)| <= 2^-11 * 2^-11 * |w| / 2 (+ fp16 rounding of the tail)
    assert err <= 2.0 ** -22, (kind, pq, err)
    m, nb, groups, gbl, blocks, pool, split, smem = list(info)
    m, a_slots = m & 0xffff, m >> 16                        # operand-ring depth rides in the high half
    assert nb in (16, 32) and m >= 1 and blocks >= 1 and groups == -(-(pq[1] * m) // nb)
    assert a_slots in (2, 4)
    assert gbl * 2 * nb + pool * nb <= 512 - 32 * a_slots  # TMEM: accumulators + pool below the operand ring
    assert smem <= 227 * 1024
    if pool:
        assert kind == 0 and split > 0                      # the accumulator split is for long windows only


def test_bench_ratio_plan(f9):
    """config 2 (96 kHz -> 44.1 kHz, WindowedSinc): 32-slot groups, one block, accumulator split on."""
    info = (C.c_int * 8)()
    assert f9.lib().f9_umma_selfcheck(0, 320, 147, info) >= 0
    m, nb, groups, gbl, blocks, pool, split, smem = list(info)
    assert (m & 0xffff, nb, groups, blocks) == (1, 32, 5, 1) and pool >= 2 and split >= 8
    assert m >> 16 == 4                                     # four-stage operand ring: the pool of two slots still fits


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
@pytest.mark.parametrize("up", [2, 4, 8, 16])
def test_hankel_weight_image(f9, kind, up):
    """Weight image of the Hankel-operand FIR (integer upsampling): lane i*L + k holds phase k's taps shifted by i, zeros elsewhere;
    the tile buffer holds every K step of every column; the kernel's shared memory and TMEM budgets hold."""
    info = (C.c_int * 4)()
    err = f9.lib().f9_hankel_selfcheck(kind, up, info)
    assert 0.0 <= err <= 2.0 ** -22, (kind, up, err)
    ks, elems, buf_bytes, smem = list(info)
    r = 128 // up
    assert ks * 16 >= r + 208 and ks in (14, 15, 17)
    assert elems == r * 64 + 16 * ks and elems % 8 == 0 and elems <= 3 * 256 * 8      # three groups of 8 per converter thread
    assert 32 * (ks - 1) + 63 * 2 * r + 32 <= 2 * elems                                 # last K step of the last column stays inside
    assert buf_bytes % 1024 == 0 and buf_bytes >= 2 * elems and smem <= 227 * 1024
    assert 8 * ks <= 136 and 272 + 2 * 96 <= 512                                        # TMEM: weight heads, tails, two accumulator sets


def test_hankel_bad_arguments(f9):
    assert f9.lib().f9_hankel_selfcheck(0, 3, None) == -1.0
    assert f9.lib().f9_hankel_selfcheck(99, 4, None) == -1.0


def test_bad_arguments(f9):
    assert f9.lib().f9_umma_selfcheck(99, 1, 1, None) == -1.0
    assert f9.lib().f9_umma_selfcheck(0, 0, 1, None) == -1.0
