"""``mindaudio.data.augment`` pieces on the feature path: spectrogram masking, time_stretch (phase vocoder), pitch_shift."""
from mindaudio_b200.data.augment import _phase_vocoder, frequencymasking, pitch_shift, time_stretch, timemasking  # noqa: F401
