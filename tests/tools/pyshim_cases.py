"""Runs the cases of the reference's own binding tests (pygenomeworks/test/test_cudapoa_bindings.py,
test_cudaaligner_bindings.py) through the reference's Cython shim built on THIS engine (oracle/build_pyshim.py ->
oracle/_ref/pygw). Executed in its own process by tests/test_gpu_pyshim.py (the shim's package is called `genomeworks`,
like this repo's import alias). Prints PYSHIM_OK on success."""
import os
import random
import sys
from difflib import SequenceMatcher

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "pygw"))

import genomeworks.cuda as cuda  # noqa: E402
from genomeworks.cudaaligner import CudaAlignerBatch  # noqa: E402
from genomeworks.cudapoa import CudaPoaBatch  # noqa: E402
import genomeworks.cudapoa.cudapoa as _poa_mod  # noqa: E402

assert "oracle/_ref/pygw" in _poa_mod.__file__.replace("\\", "/"), _poa_mod.__file__

device = cuda.cuda_get_device()
free, total = cuda.cuda_get_mem_info(device)
mem = min(0.9 * free, 8 * 2**30)

# test_cudapoa_simple_batch (test_cudapoa_bindings.py:27-44)
batch = CudaPoaBatch(10, 1024, mem, deivce_id=device, output_mask='consensus')
batch.add_poa_group(["ACTGACTG", "ACTTACTG", "ACGGACTG", "ATCGACTG"])
batch.add_poa_group(["ACTGAC", "ACTTAC", "ACGGAC", "ATCGAC"])
batch.generate_poa()
consensus, coverage, status = batch.get_consensus()
assert len(consensus) == 2 and batch.total_poas == 2, (consensus, status)
assert consensus[0] == "ACTGACTG" and consensus[1] == "ACTGAC", consensus
del batch

# test_cudapoa_banded_aligned_batch (:47-64)
batch = CudaPoaBatch(10, 1024, mem, deivce_id=device, output_mask='consensus', cuda_banded_alignment=True)
batch.add_poa_group(["ACTGACTG", "ACTTACTG", "ACGGACTG", "ATCGACTG"])
batch.add_poa_group(["ACTGAC", "ACTTAC", "ACGGAC", "ATCGAC"])
batch.generate_poa()
consensus, coverage, status = batch.get_consensus()
assert len(consensus) == 2 and batch.total_poas == 2
del batch

# test_cudapoa_incorrect_output_type / valid_output_type (:67-89)
try:
    CudaPoaBatch(10, 1024, mem, deivce_id=device, output_type='error_input')
    raise AssertionError("no RuntimeError for a bad output type")
except RuntimeError:
    pass
CudaPoaBatch(10, 1024, mem, deivce_id=device, output_type='consensus')

# test_cudapoa_reset_batch (:92-106)
batch = CudaPoaBatch(10, 1024, mem, device_id=device)
batch.add_poa_group(["ACTGACTG", "ACTTACTG", "ACGGACTG", "ATCGACTG"])
batch.generate_poa()
batch.get_consensus()
assert batch.total_poas == 1
batch.reset()
assert batch.total_poas == 0
del batch

# test_cudapoa_graph (:109-134): 10 nodes / 11 edges
batch = CudaPoaBatch(10, 1024, mem, device_id=device)
batch.add_poa_group(["ACTGACTG", "ACTTACTG", "ACTCACTG"])
batch.generate_poa()
batch.get_consensus()
graphs, status = batch.get_graphs()
assert len(graphs) == 1
assert graphs[0].number_of_nodes() == 10 and graphs[0].number_of_edges() == 11
del batch

# test_cudapoa_complex_batch (:137-162): 100 reads x 500 bp @ 2 % -> consensus == reference
random.seed(2)
read_len = 500
ref = ''.join([random.choice(['A', 'C', 'G', 'T']) for _ in range(read_len)])
reads = []
for _ in range(100):
    reads.append(''.join([r if random.random() > 0.02 else random.choice(['A', 'C', 'G', 'T']) for r in ref]))
batch = CudaPoaBatch(1000, 1024, mem, device_id=device)
add_status, seq_status = batch.add_poa_group(reads)
batch.generate_poa()
consensus, coverage, status = batch.get_consensus()
assert len(consensus[0]) == len(ref)
assert SequenceMatcher(None, ref, consensus[0]).ratio() == 1.0
del batch

# MSA through the shim
batch = CudaPoaBatch(10, 1024, mem, device_id=device, output_type='msa')
batch.add_poa_group(["ACTGACTG", "ACTTACTG", "ACTCACTG"])
batch.generate_poa()
msa, status = batch.get_msa()
assert len(msa) == 1 and len(msa[0]) == 3 and len(set(len(r) for r in msa[0])) == 1, msa
del batch

# test_cudaaligner_simple_batch (test_cudaaligner_bindings.py:27-45)
for query, target, cigar in [("AAAAAAA", "TTTTTTT", "7M"), ("AAATC", "TACGTTTT", "3M1I2M2I"), ("TACGTA", "ACATAC", "1D5M1I"),
                             ("TGCA", "ATACGCT", "1I1M2I3M")]:
    stream = cuda.CudaStream()
    ab = CudaAlignerBatch(len(query), len(target), 1, alignment_type="global", stream=stream, device_id=device)
    ab.add_alignment(query, target)
    ab.align_all()
    alignments = ab.get_alignments()
    assert len(alignments) == 1
    assert alignments[0].cigar == cigar, (query, target, alignments[0].cigar, cigar)
    del ab

# test_cudaaligner_long_alignments (:48-74), own generator in place of the pure-Python simulators
rng = random.Random(7)
for ref_length, num_alignments in [(5000, 30), (10000, 10), (500, 100)]:
    ab = CudaAlignerBatch(ref_length, ref_length, num_alignments, device_id=device)
    for _ in range(num_alignments):
        reference = ''.join(rng.choice("ACGT") for _ in range(ref_length))
        q = ''.join(c for c in reference if rng.random() > 0.02)
        t = ''.join(c if rng.random() > 0.03 else rng.choice("ACGT") for c in reference)
        assert ab.add_alignment(q, t) == 0
    ab.align_all()
    al = ab.get_alignments()
    assert len(al) == num_alignments
    ab.reset()
    assert len(ab.get_alignments()) == 0
    del ab

# test_cudaaligner_various_arguments (:77-105)
for max_seq_len, max_alignments, seq_len, num_alignments, should_succeed in [(1000, 100, 10000, 10, False), (1000, 100, 100, 10, True),
                                                                             (1000, 100, 1000, 100, True), (100, 10, 100, 1000, False)]:
    ab = CudaAlignerBatch(max_seq_len, max_seq_len, max_alignments, device_id=device)
    success = True
    for _ in range(num_alignments):
        reference = ''.join(rng.choice("ACGT") for _ in range(seq_len))
        q = ''.join(c for c in reference if rng.random() > 0.02)
        t = ''.join(c for c in reference if rng.random() > 0.02)
        if ab.add_alignment(q, t) != 0:
            success = False
    ab.align_all()
    assert success is should_succeed, (max_seq_len, max_alignments, seq_len, num_alignments)
    del ab

print("PYSHIM_OK")
