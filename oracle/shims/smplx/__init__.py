"""Import shim: `smplx` is absent from this image and only needed by the reference's `lib/models/smpl_mps.py` (imported by
`lib/models/spin.py:12`) for the HMR regression HEAD, which is not on the feature-extractor path (spin.py:129-143). TEST INFRASTRUCTURE."""


class SMPL:            # smplx.SMPL: never instantiated by the oracle harness (HMR.__init__ is bypassed)
    def __init__(self, *a, **k):
        raise RuntimeError("smplx shim: the SMPL body model is not available in this image")
