"""gw-b200: Blackwell-native batched partial-order alignment and banded pairwise alignment.

Drop-in for the cudapoa::Batch and cudaaligner::Aligner paths of GenomeWorks behind a C ABI (include/gwb200.h).
Python surface mirrors pygenomeworks: genomeworks_b200.cudapoa.CudaPoaBatch, genomeworks_b200.cudaaligner.CudaAlignerBatch,
genomeworks_b200.cuda.CudaStream. No CPU fallback: the in-tree CUDA library must be built and a GPU present.
"""
__version__ = "0.1.0"
