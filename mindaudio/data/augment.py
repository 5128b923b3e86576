"""``mindaudio.data.augment`` pieces on the feature path: spectrogram masking, time_stretch (phase vocoder), pitch_shift,
speed_perturb."""
from mindaudio_b200.data.augment import (_phase_vocoder, frequencymasking, pitch_shift, speed_perturb,  # noqa: F401
                                         time_stretch, timemasking)
