"""Process-wide side streams, one per ROLE and device.

torch hands out `torch.cuda.Stream()` objects from a pool of 32 per device, round-robin: the 33rd stream a process creates is
the first one again.  Round 1 created a fresh stream per module instance (position-embedding stream, text stream, label stream,
two value-projection streams, capture stream, ...), so a process that had built a few models before (a test suite, a sweep)
ended up with e.g. the capture stream being the same CUDA stream as a criterion branch stream and the text stream the same as
another branch stream.  With that aliasing the graphed train step produced non-finite parameters during its capture (r02s:
reproduced and the handles printed by tools/debug_nan_sequence.py).  Every side stream of the package now comes from this
registry: one stream per (device, role), created once - at most ~16 per device, never recycled, never aliased, and independent of
how many models a process builds (two model instances share the role streams; they do not run concurrently)."""
import torch

_STREAMS = {}


def get(device, role):
    """the stream of `role` (a short string, e.g. 'text', 'lang', 'value0', 'branch3') on `device`"""
    device = torch.device(device) if not isinstance(device, torch.device) else device
    if device.type != "cuda":
        raise ValueError("side streams exist on CUDA devices only")
    index = device.index if device.index is not None else torch.cuda.current_device()
    key = (index, role)
    st = _STREAMS.get(key)
    if st is None:
        assert len([k for k in _STREAMS if k[0] == index]) < 28, "role streams would start to alias torch's pool of 32"
        st = _STREAMS[key] = torch.cuda.Stream(torch.device("cuda", index))
    return st


def handles(device=None):
    """{role: cudaStream_t} of the streams created so far (diagnostics / tests)"""
    return {role: st.cuda_stream for (idx, role), st in _STREAMS.items()
            if device is None or idx == torch.device(device).index}
