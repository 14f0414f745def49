"""TensorFlow-free restatement of the hot-path hyper-parameters.

The reference keeps them in ``tf.contrib.training.HParams`` singletons
(``tacotron/params/model.py:7-48``, ``tacotron/params/inference.py:34``); only the values the
audio path reads are restated here, with identical names.
"""


class _ModelParams:
    vocabulary_size = 39
    sampling_rate = 22050               # tacotron/params/model.py:13
    n_fft = 2048                        # :16
    win_len = 50.0                      # :20  (ms)
    win_hop = 12.5                      # :24  (ms)
    n_mels = 80                         # :27
    mel_fmin = 0                        # :30
    mel_fmax = 8000                     # :33
    n_mfcc = 13                         # :36
    reduction = 5                       # :39
    magnitude_power = 1.3               # :45
    reconstruction_iterations = 50      # :48


model_params = _ModelParams()


class _InferenceParams:
    n_synthesis_threads = 6             # tacotron/params/inference.py:34 (unused: one batched GPU call)


inference_params = _InferenceParams()
