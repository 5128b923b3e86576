"""``mindaudio.data.augment`` pieces on the feature path: spectrogram masking, time_stretch (phase vocoder)."""
from mindaudio_b200.data.augment import _phase_vocoder, frequencymasking, time_stretch, timemasking  # noqa: F401
