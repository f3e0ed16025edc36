"""Mirror of pygenomeworks' genomeworks.cuda (pygenomeworks/genomeworks/cuda/cuda.pyx:28-135): CudaRuntimeError, CudaStream,
cuda_get_device_count / cuda_set_device / cuda_get_device / cuda_get_mem_info, implemented over torch.cuda (device memory and
stream plumbing only)."""
import torch


class CudaRuntimeError(Exception):
    """Raised when a CUDA runtime call fails (cuda.pyx:28-35)."""

    def __init__(self, error):
        super().__init__("CUDA runtime error: %s" % (error,))


def _require_cuda():
    if not torch.cuda.is_available():
        raise CudaRuntimeError("no CUDA device (genomeworks_b200 has no CPU fallback)")


class CudaStream:
    """RAII CUDA stream (cuda.pyx:38-80). `.stream` is the raw cudaStream_t value."""

    def __init__(self, device=None):
        _require_cuda()
        self._stream = torch.cuda.Stream(device=device)

    @property
    def stream(self):
        return self._stream.cuda_stream

    def sync(self):
        self._stream.synchronize()


def cuda_get_device_count():
    """Number of CUDA-capable devices (cuda.pyx:83-93)."""
    return torch.cuda.device_count()


def cuda_set_device(device_id):
    """Makes `device_id` the current device (cuda.pyx:96-101)."""
    _require_cuda()
    try:
        torch.cuda.set_device(device_id)
    except Exception as e:  # invalid ordinal etc.
        raise CudaRuntimeError(e)


def cuda_get_device():
    """Current device id (cuda.pyx:104-114)."""
    _require_cuda()
    return torch.cuda.current_device()


def cuda_get_mem_info(device_id):
    """Returns (free, total) bytes of device `device_id` (cuda.pyx:117-135)."""
    _require_cuda()
    return torch.cuda.mem_get_info(device_id)
