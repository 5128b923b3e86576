"""CPU oracle for the front-end feature path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg may import this package.  Nothing under
``mindaudio_b200/`` imports it; the product path fails loudly when the CUDA
library is missing instead of falling back to this code.

Layout
------
``restated.py``   float64 numpy restatement of every function on the path
                  (SURVEY.md section 8a); each function cites the reference
                  file:line it follows.  Travels to the GPU box.
``ref_loader.py`` loads the reference's OWN python (``/root/reference``)
                  unchanged, around a stub ``mindspore`` package
                  (``ms_shim/``).  Only usable in the dev container.
``make_goldens.py`` runs the reference code and freezes ``tests/golden/*.npz``.
``make_wav_goldens.py`` does the same for ``mindaudio.data.io.read`` over the WAV
                  corpus of ``tests/wav_util.py`` (``tests/golden/wav_io.npz``).

Parity status (see DESIGN.md section "Oracle"):
* a1-a5, a10-a13 (stft/istft/magphase/dB, conformer fbank, CMVN): PINNED --
  ``restated.py`` is checked against the reference's own code executed here
  and against the frozen goldens.
* f1-f4 rows: ``context_window``, ``pad_sequence`` / ``make_pad_mask``, ``io.read`` (27 container
  cases + 12 000 fuzzed mutations), conformer ``spec_aug``, ``_phase_vocoder`` / ``time_stretch`` /
  ``resample`` / ``pitch_shift``, ``soft_mask`` / ``hpss`` / ``harmonic``: PINNED the same way.
* a6-a9 (spectrogram/melspectrogram/melscale/fbank/mfcc): the arithmetic lives
  in mindspore==2.3.0 C++ (``mindspore.dataset.audio``), which is neither
  vendored under /root/reference nor installable here.  Restated from the
  published API contract (torchaudio-lineage formulas) and cross-checked
  against ``torchaudio.functional``; **parity unpinned vs the MindSpore
  binary**.  The reference's own python wrapped around the restated ops IS
  executed for the goldens, so defaults / hop rules / dB / top_db / transposes
  are the reference's.
"""
