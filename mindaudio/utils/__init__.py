"""Alias of the reference's ``mindaudio.utils`` for the pieces on the feature path (collate helpers, CMVN loading)."""
