"""CPU: the oracle's restatement of juce::ResamplingAudioSource -- parity unpinned (JUCE is
not vendored), so these pin the properties the algorithm must have whatever the platform."""
import numpy as np
import pytest




def noise(n, seed, ch=1):
    return np.random.default_rng(seed).uniform(-0.5, 0.5, (ch, n)).astype(np.float32)


@pytest.mark.parametrize("ratio", [320 / 147, 0.25, 147 / 160, 1.0, 3.3333])
def test_block_size_does_not_change_the_samples(O, ratio):
    x = noise(30000, 1, 2)
    n_out = int(30000 / ratio) - 8
    a = O.ras_convert(x, ratio, n_out, 512)
    for block in (1, 77, 4096):
        assert np.array_equal(a, O.ras_convert(x, ratio, n_out, block)), block


def test_unity_ratio_is_a_copy_and_dc_gain_is_one(O):
    x = noise(5000, 2)
    assert np.array_equal(O.ras_convert(x, 1.0, 4000), x[:, :4000])
    dc = np.full((1, 6000), 0.25, dtype=np.float32)
    assert abs(float(O.ras_convert(dc, 2.0, 2500)[0, -1]) - 0.25) < 1e-6          # pre-filter, settled
    assert abs(float(O.ras_convert(dc, 0.5, 9000)[0, -1]) - 0.25) < 1e-6          # post-filter, settled


def test_pull_counts_follow_juce(O):
    """getNextAudioBlock pulls round(numSamples * ratio) + 3 minus what is still buffered."""
    x = noise(100000, 3)
    src = O.ResamplingAudioSource(x)
    ratio = 320 / 147
    src.set_resampling_ratio(ratio)
    src.prepare_to_play(512)
    have, sub = 0, 0.0
    for n in (512, 512, 100, 1, 2048):
        src.get_next_audio_block(n)
        need = int(np.rint(n * ratio)) + 3
        assert src.pulled == max(0, need - have)
        have += src.pulled
        for _ in range(n):
            sub += ratio
            while sub >= 1.0:
                sub -= 1.0
                have -= 1


def test_coefficients_are_a_unit_dc_gain_butterworth(O):
    for ratio in (2.0, 0.5, 320 / 147, 0.25):
        c = O.resampling_source_coeffs(ratio)
        assert abs((c[0] + c[1] + c[2]) / (1.0 + c[4] + c[5]) - 1.0) < 1e-12       # H(1) = 1
        disc = c[4] * c[4] - 4 * c[5]
        assert disc < 0 and 0 < c[5] < 1                                            # complex pole pair inside the unit circle
