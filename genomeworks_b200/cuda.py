"""Mirror of pygenomeworks' genomeworks.cuda (pygenomeworks/genomeworks/cuda/cuda.pyx:38-117): CudaStream and
cuda_get_mem_info, implemented over torch.cuda (device memory / stream plumbing only)."""
import torch


class CudaStream:
    """RAII CUDA stream (cuda.pyx CudaStream). `.stream` is the raw cudaStream_t value."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA device required (genomeworks_b200 has no CPU fallback)")
        self._stream = torch.cuda.Stream(device=device)

    @property
    def stream(self):
        return self._stream.cuda_stream

    def sync(self):
        self._stream.synchronize()


def cuda_get_mem_info(device_id):
    """Returns (free, total) bytes of device `device_id` (cuda.pyx cuda_get_mem_info)."""
    if not torch.cuda.is_available():
        raise RuntimeError("CUDA device required (genomeworks_b200 has no CPU fallback)")
    return torch.cuda.mem_get_info(device_id)
