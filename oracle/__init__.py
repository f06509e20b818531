"""CPU oracle for the loe_speech_recognition hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as the thing shipped.  The
product path (``cs-304-speech-recognition-code_b200/``) never imports this
package and fails loudly when its CUDA library is missing.

Contents
--------
``mfcc.py``      restated librosa MFCC pipeline used by the reference
                 (``src/loe_speech_recognition/mfcc.py:24-69``).  librosa is an
                 un-vendored, un-pinned third-party dependency that is absent
                 from this image and the reference holds no golden vectors for
                 it, so this part is **parity unpinned** (SURVEY.md §8c).
``hmm.py``       vectorised NumPy restatement of emission scoring, word / loop
                 / chain Viterbi, label decoding, the segmental K-means M-step
                 and the embedded-training remux
                 (``src/loe_speech_recognition/hidden_markov_model.py``,
                 ``signal.py``, ``model_boundary.py``,
                 ``transition_probability.py``).  Pinned bit-for-bit against the
                 unmodified reference run in the authoring container
                 (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).
``ref_port.py``  structure-faithful port of the reference's per-(frame,state)
                 Python/scipy loops; this is what ``bench.py`` times as the
                 reference CPU path (``cpu_baseline.kind == "port"``).
``ref_import.py`` imports the *real* reference from ``/root/reference`` with
                 stub modules for its missing GUI/audio dependencies.  Only
                 usable in the authoring container; never touched on the GPU box.
"""
