"""GPU: juce::ResamplingAudioSource (f9_ras_*) against the oracle's restatement.
Streaming object: same control flow on the host, JUCE's operation order on the device -> bit-exact blocks and pull counts.
Whole-channel batch: chunk-parallel IIR + closed-form position -> tolerance (max-abs 2^-20, SNR >= 120 dB)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 2.0 ** -20


def noise(n, seed, ch=2):
    return np.random.default_rng(seed).uniform(-0.5, 0.5, (ch, n)).astype(np.float32)


def snr_db(ref, got):
    e = np.sqrt(np.mean((ref.astype(np.float64) - got.astype(np.float64)) ** 2))
    s = np.sqrt(np.mean(ref.astype(np.float64) ** 2))
    return np.inf if e == 0 else 20 * np.log10(s / e)


@pytest.mark.parametrize("ratio", [320 / 147, 0.25, 147 / 160, 1.0, 1.00005, 3.7, 0.731234567])
def test_streaming_object_is_bit_exact(ctx, O, f9, ratio):
    x = noise(60000, 11)
    a, b = f9.ResamplingAudioSource(ctx, x), O.ResamplingAudioSource(x)
    for s in (a, b):
        s.set_resampling_ratio(ratio)
        s.prepare_to_play(512)
    rng = np.random.default_rng(12)
    for _ in range(12):
        n = int(rng.integers(1, 1500))
        ya, yb = a.get_next_audio_block(n), b.get_next_audio_block(n)
        assert a.pulled == b.pulled
        assert np.array_equal(ya, yb), (ratio, n)


def test_streaming_ratio_change_and_flush(ctx, O, f9):
    x = noise(80000, 13, ch=3)
    a, b = f9.ResamplingAudioSource(ctx, x), O.ResamplingAudioSource(x)
    for s in (a, b):
        s.set_resampling_ratio(2.0)
        s.prepare_to_play(256)
    for ratio, n in ((2.0, 700), (0.5, 900), (1.0, 300), (1.0, 1), (2.5, 1000), (0.3, 2000)):
        for s in (a, b):
            s.set_resampling_ratio(ratio)
        ya, yb = a.get_next_audio_block(n), b.get_next_audio_block(n)
        assert a.pulled == b.pulled and np.array_equal(ya, yb), (ratio, n)
    for s in (a, b):
        s.flush_buffers()
    assert np.array_equal(a.get_next_audio_block(512), b.get_next_audio_block(512))


def test_streaming_past_the_end_of_the_source_reads_zeros(ctx, O, f9):
    x = noise(3000, 14, ch=1)
    a, b = f9.ResamplingAudioSource(ctx, x), O.ResamplingAudioSource(x)
    for s in (a, b):
        s.set_resampling_ratio(320 / 147)
        s.prepare_to_play(1024)
    for _ in range(4):
        assert np.array_equal(a.get_next_audio_block(1024), b.get_next_audio_block(1024))


@pytest.mark.parametrize("ratio", [320 / 147, 0.25, 147 / 160, 2.0, 1.0, 0.731234567, 1.001])
def test_whole_channels_within_tolerance(ctx, O, ratio):
    x = noise(300000, 15, ch=4)
    n_out = int(300000 / ratio) - 8
    y = ctx.ras_convert(x, ratio, n_out)
    ref = O.ras_convert(x, ratio, n_out, 4096)
    assert np.max(np.abs(y - ref)) <= TOL, ratio
    assert snr_db(ref, y) >= 120.0, ratio


def test_whole_channels_output_past_the_input(ctx, O):
    """Outputs that read past the end of the file see the filter ringing on zeros, as a streaming run would."""
    x = noise(20000, 16, ch=2)
    for ratio in (320 / 147, 0.5):
        n_out = int(20000 / ratio) + 600
        y, ref = ctx.ras_convert(x, ratio, n_out), O.ras_convert(x, ratio, n_out, 512)
        assert np.max(np.abs(y - ref)) <= TOL, ratio


def test_device_batch_entry_point(ctx, O, f9):
    torch = pytest.importorskip("torch")
    x = noise(200000, 17, ch=8)
    ratio = 96000 / 44100
    n_out = int(200000 / ratio) - 8
    d_in = torch.from_numpy(x).cuda()
    d_out = torch.zeros((8, n_out), dtype=torch.float32, device="cuda")
    sf = f9.lib().f9_ras_scratch_frames(ratio, n_out)
    d_scr = torch.empty((8, sf), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    assert f9.lib().f9_dev_ras_convert(ctx.handle, d_in.data_ptr(), 200000, 8, 200000, ratio, d_out.data_ptr(), n_out, n_out,
                                       d_scr.data_ptr(), sf) == 0
    ctx.synchronize()
    ref = O.ras_convert(x, ratio, n_out, 4096)
    assert np.max(np.abs(d_out.cpu().numpy() - ref)) <= TOL
