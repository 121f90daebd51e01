"""Import-time compatibility for the (2020) reference code base under numpy >= 1.24 / Python >= 3.9.

The reference is not edited; two attribute aliases it expects are restored in this process (the same two
shims the test oracle loader applies):

* ``apply_numpy()`` -- ``np.float`` / ``np.int`` / ``np.bool`` (fairseq/data/indexed_dataset.py:89): needed
  BEFORE ``import fairseq``, i.e. before fairseq can import any ``--user-dir``.  ``run_fairseq.py`` in this
  directory is a launcher that does it and then hands over to fairseq's own ``train`` / ``generate`` CLI;
* ``apply_fairseq()`` -- ``EncoderOut._field_types`` (examples/speech_recognition/models/conv_transformer.py:30
  reads the ``NamedTuple`` attribute removed in Python 3.9): applied by the plugin itself.
"""


def apply_numpy():
    import numpy as np
    for name, typ in (("float", float), ("int", int), ("bool", bool)):
        if name not in np.__dict__:
            setattr(np, name, typ)


def apply_fairseq():
    from fairseq.models import fairseq_encoder as fe
    if not hasattr(fe.EncoderOut, "_field_types"):
        fe.EncoderOut._field_types = fe.EncoderOut.__annotations__


def apply():
    apply_numpy()
    apply_fairseq()
