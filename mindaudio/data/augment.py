"""``mindaudio.data.augment`` pieces on the feature path: frequency / time masking of spectrograms."""
from mindaudio_b200.data.masking import frequencymasking, timemasking  # noqa: F401
