"""Python mirror of pygenomeworks' cudaaligner shim (pygenomeworks/genomeworks/cudaaligner/cudaaligner.pyx) over the gw-b200
C ABI, plus the FixedBandAligner surface (cudaaligner/include/.../aligner.hpp:148-219) the reference's Python shim lacks.

CudaAlignerBatch keeps the deprecated-factory signature (max_query_length, max_target_length, max_alignments, ...). The
reference serves it with AlignerGlobalHirschbergMyers; this engine serves it with the banded Myers aligner at a bandwidth
that covers the whole query (full Myers, exact), whose tie-breaking (insertion, deletion, diagonal) reproduces the CIGARs the
reference's Python tests pin (tests/test_oracle_aligner.py)."""
import ctypes as C

import numpy as np

from ._lib import check, lib

success = 0
uninitialized = 1
exceeded_max_alignments = 2
exceeded_max_length = 3
exceeded_max_alignment_difference = 4
generic_error = 5

match, mismatch, insertion, deletion = 0, 1, 2, 3
global_alignment = 0

_STATUS = ["success", "uninitialized", "exceeded_max_alignments", "exceeded_max_length", "exceeded_max_alignment_difference",
           "generic_error"]
_BASIC = {0: "M", 1: "M", 2: "I", 3: "D"}
_EXT = {0: "=", 1: "X", 2: "I", 3: "D"}


def status_to_str(status):
    if 0 <= int(status) < len(_STATUS):
        return _STATUS[int(status)]
    raise RuntimeError("Unknown error status : " + str(status))


def _cigar(actions, runs, extended=False):
    """AlignmentImpl::convert_to_cigar over the RLE form (cudaaligner/src/alignment_impl.cpp:99-153)."""
    if extended:
        return "".join("%d%s" % (int(r), _EXT[int(a)]) for a, r in zip(actions, runs))
    out, last, cnt = [], None, 0
    for a, r in zip(actions, runs):
        c = _BASIC[int(a)]
        if c == last:
            cnt += int(r)
        else:
            if last is not None:
                out.append("%d%s" % (cnt, last))
            last, cnt = c, int(r)
    if last is not None:
        out.append("%d%s" % (cnt, last))
    return "".join(out)


_RC_LOOKUP = b"TGAC"


def _revcomp(seq):
    """genomeutils::reverse_complement (utils/genomeutils.hpp:144-154): lookup by (c >> 1) & 3 over the reversed sequence."""
    return bytes(_RC_LOOKUP[(c >> 1) & 3] for c in reversed(seq))


class Alignment:
    """What cudaaligner::Alignment exposes (alignment.hpp:55-111)."""

    def __init__(self, query, target, status, is_optimal, actions, runs):
        self.query = query
        self.target = target
        self.status = status
        self.is_optimal = bool(is_optimal)
        self.actions = actions
        self.runlengths = runs

    def get_query_sequence(self):
        return self.query

    def get_target_sequence(self):
        return self.target

    def convert_to_cigar(self, extended=False):
        return _cigar(self.actions, self.runlengths, extended)

    def get_edit_distance(self):
        return int(sum(int(r) for a, r in zip(self.actions, self.runlengths) if int(a) != match))

    def get_alignment(self):
        out = []
        for a, r in zip(self.actions, self.runlengths):
            out.extend([int(a)] * int(r))
        return out

    def format_alignment(self):
        """(query line, pairing line, target line), cf. AlignmentImpl::format_alignment."""
        q, p, t = [], [], []
        qi = ti = 0
        for s in self.get_alignment():
            if s in (match, mismatch):
                q.append(self.query[qi])
                t.append(self.target[ti])
                p.append("|" if s == match else "x")
                qi += 1
                ti += 1
            elif s == insertion:
                q.append("-")
                t.append(self.target[ti])
                p.append(" ")
                ti += 1
            else:
                q.append(self.query[qi])
                t.append("-")
                p.append(" ")
                qi += 1
        return ("".join(q), "".join(p), "".join(t))


class CudaAlignment:
    """pygenomeworks' result record (cudaaligner.pyx:57-126)."""

    def __init__(self, query, target, cigar, alignment_type, status, alignment, format_alignment):
        self.query = query
        self.target = target
        self.cigar = cigar
        self.alignment_type = "global"
        self.status = status
        names = {match: "m", mismatch: "mm", insertion: "i", deletion: "d"}
        self.alignment = [names[s] for s in alignment]
        self.format_alignment = format_alignment

    def __str__(self):
        return "{}\n{}\n{}\n".format(self.format_alignment[0], self.format_alignment[1], self.format_alignment[2])


class FixedBandAligner:
    """create_aligner(AlignmentType::global_alignment, max_bandwidth, stream, device_id, max_device_memory) -- aligner.hpp:208-219."""

    def __init__(self, max_bandwidth, stream=None, device_id=0, max_device_memory=-1):
        self._h = C.c_void_p()
        st = None
        if stream is not None:
            st = stream.stream if hasattr(stream, "stream") else stream.cuda_stream
        self.stream = stream
        self._pairs = []
        self._results = []
        check(lib().gwb200_aligner_create(C.byref(self._h), C.c_int32(max_bandwidth), C.c_void_p(st), C.c_int32(device_id),
                                          C.c_int64(int(max_device_memory))))

    def add_alignment(self, query, target, max_bandwidth=None, reverse_complement_query=False, reverse_complement_target=False):
        if max_bandwidth is None:
            max_bandwidth = -2147483648  # GWB200_ALN_DEFAULT_BANDWIDTH: the aligner's own
        q = query.encode("utf-8") if isinstance(query, str) else bytes(query)
        t = target.encode("utf-8") if isinstance(target, str) else bytes(target)
        rc = check(lib().gwb200_aligner_add_alignment(self._h, C.c_int32(max_bandwidth), q, C.c_int32(len(q)), t, C.c_int32(len(t)),
                                                      C.c_int32(1 if reverse_complement_query else 0),
                                                      C.c_int32(1 if reverse_complement_target else 0)))
        if rc == success:
            # the Alignment carries the sequences as aligned (after the reverse complement the flags ask for), as the reference
            # builds it from its staging copy (aligner_global_myers_banded.cpp:226-227,413-418)
            self._pairs.append((_revcomp(q) if reverse_complement_query else q, _revcomp(t) if reverse_complement_target else t))
        return rc

    def add_alignments(self, pairs):
        """add_alignment() for a list of (query, target) byte strings in one call into the engine (the loop a C++ caller writes).
        Returns (status of the last call, number of pairs added)."""
        n = len(pairs)
        qs = (C.c_char_p * n)(*[p[0] for p in pairs])
        ts = (C.c_char_p * n)(*[p[1] for p in pairs])
        ql = (C.c_int32 * n)(*[len(p[0]) for p in pairs])
        tl = (C.c_int32 * n)(*[len(p[1]) for p in pairs])
        added = C.c_int32(0)
        rc = check(lib().gwb200_aligner_add_alignments(self._h, C.c_int32(n), qs, ql, ts, tl, C.byref(added)))
        self._pairs.extend(pairs[:added.value])
        return rc, added.value

    def align_all(self):
        return check(lib().gwb200_aligner_align_all(self._h))

    def sync_alignments_flat(self):
        """sync_alignments() with all results fetched in one call: returns (status, is_optimal, run_offsets, actions, runlengths)
        numpy arrays; get_alignments() stays empty."""
        rc = check(lib().gwb200_aligner_sync_alignments(self._h))
        n = len(self._pairs)
        cap = sum(len(q) + len(t) for q, t in self._pairs) + 16
        st = np.zeros(n, dtype=np.int32)
        opt = np.zeros(n, dtype=np.int32)
        offs = np.zeros(n + 1, dtype=np.int64)
        act = np.zeros(cap, dtype=np.int8)
        runs = np.zeros(cap, dtype=np.int32)
        check(lib().gwb200_aligner_results_flat(self._h, st.ctypes.data, opt.ctypes.data, offs.ctypes.data, act.ctypes.data, runs.ctypes.data,
                                                C.c_int64(cap)))
        self._pairs = []
        self._results = []
        return st, opt, offs, act[:offs[n]], runs[:offs[n]]

    def sync_alignments(self, want_strings=True):
        rc = check(lib().gwb200_aligner_sync_alignments(self._h))
        self._results = []
        st, opt, nr = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        for i, (q, t) in enumerate(self._pairs):
            check(lib().gwb200_aligner_result_info(self._h, C.c_int32(i), C.byref(st), C.byref(opt), C.byref(nr)))
            a = np.zeros(max(nr.value, 1), dtype=np.int8)
            r = np.zeros(max(nr.value, 1), dtype=np.int32)
            check(lib().gwb200_aligner_result_runs(self._h, C.c_int32(i), a.ctypes.data, r.ctypes.data))
            self._results.append(Alignment(q.decode() if want_strings else q, t.decode() if want_strings else t, st.value, opt.value,
                                           a[:nr.value].copy(), r[:nr.value].copy()))
        self._pairs = []
        return rc

    def get_alignments(self):
        return list(self._results)

    def num_alignments(self):
        return lib().gwb200_aligner_num_alignments(self._h)

    def reset(self):
        self._pairs = []
        self._results = []
        check(lib().gwb200_aligner_reset(self._h))

    def reset_max_bandwidth(self, max_bandwidth):
        self._pairs = []
        check(lib().gwb200_aligner_reset_max_bandwidth(self._h, C.c_int32(max_bandwidth)))

    def free_temporary_device_buffers(self):
        check(lib().gwb200_aligner_free_temporary_device_buffers(self._h))

    def last_cells(self):
        return int(lib().gwb200_aligner_last_cells(self._h))

    def last_kernel_ms(self):
        return float(lib().gwb200_aligner_last_kernel_ms(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().gwb200_aligner_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _rle(states):
    """expanded AlignmentState bytes -> (actions, runlengths)"""
    s = np.asarray(states, dtype=np.int8)
    if s.size == 0:
        return np.zeros(0, np.int8), np.zeros(0, np.int32)
    cut = np.flatnonzero(np.diff(s)) + 1
    starts = np.concatenate(([0], cut))
    ends = np.concatenate((cut, [s.size]))
    return s[starts].copy(), (ends - starts).astype(np.int32)


class GlobalAligner:
    """The fixed-size global aligners behind the deprecated factory create_aligner(max_query_length, max_target_length,
    max_alignments, ...) (aligner.hpp:183,196 -> AlignerGlobalHirschbergMyers, cudaaligner/src/aligner.cpp:31-74) and the in-library
    unbanded AlignerGlobalMyers. Host semantics = AlignerGlobal (cudaaligner/src/aligner_global.cpp:50-197)."""

    ALGORITHMS = {"hirschberg_myers": 0, "myers": 1, "ukkonen": 2}

    def __init__(self, max_query_length, max_target_length, max_alignments, algorithm="hirschberg_myers", stream=None, device_id=0):
        self._h = C.c_void_p()
        st = None
        if stream is not None:
            st = stream.stream if hasattr(stream, "stream") else stream.cuda_stream
        self.stream = stream
        self._alignments = []
        self._first_unsynced = 0
        check(lib().gwb200_global_aligner_create(C.byref(self._h), C.c_int32(self.ALGORITHMS[algorithm]), C.c_int32(max_query_length),
                                                 C.c_int32(max_target_length), C.c_int32(max_alignments), C.c_void_p(st), C.c_int32(device_id),
                                                 None, None, None))

    def add_alignment(self, query, target, reverse_complement_query=False, reverse_complement_target=False):
        q = query.encode("utf-8") if isinstance(query, str) else bytes(query)
        t = target.encode("utf-8") if isinstance(target, str) else bytes(target)
        rc = check(lib().gwb200_global_aligner_add_alignment(self._h, q, C.c_int32(len(q)), t, C.c_int32(len(t)),
                                                             C.c_int32(1 if reverse_complement_query else 0),
                                                             C.c_int32(1 if reverse_complement_target else 0)))
        if rc == success:
            # the Alignment exists from here on, status uninitialized until sync_alignments (aligner_global.cpp:131-138)
            qa = _revcomp(q) if reverse_complement_query else q
            ta = _revcomp(t) if reverse_complement_target else t
            self._alignments.append(Alignment(qa.decode(), ta.decode(), uninitialized, False, np.zeros(0, np.int8), np.zeros(0, np.int32)))
        return rc

    def align_all(self):
        return check(lib().gwb200_global_aligner_align_all(self._h))

    def sync_alignments(self):
        rc = check(lib().gwb200_global_aligner_sync_alignments(self._h))
        have, opt, n = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        for i, al in enumerate(self._alignments):
            check(lib().gwb200_global_aligner_result_info(self._h, C.c_int32(i), C.byref(have), C.byref(opt), C.byref(n)))
            if not have.value:
                continue
            st = np.zeros(max(n.value, 1), dtype=np.int8)
            check(lib().gwb200_global_aligner_result_states(self._h, C.c_int32(i), st.ctypes.data))
            al.actions, al.runlengths = _rle(st[:n.value])
            al.is_optimal = bool(opt.value)
            al.status = success
        return rc

    def get_alignments(self):
        return list(self._alignments)

    def num_alignments(self):
        return len(self._alignments)

    def reset(self):
        self._alignments = []
        check(lib().gwb200_global_aligner_reset(self._h))

    def last_cells(self):
        return int(lib().gwb200_global_aligner_last_cells(self._h))

    def last_kernel_ms(self):
        return float(lib().gwb200_global_aligner_last_kernel_ms(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().gwb200_global_aligner_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CudaAlignerBatch:
    """pygenomeworks.CudaAlignerBatch (cudaaligner.pyx:129-260): deprecated-factory signature."""

    def __init__(self, max_query_length, max_target_length, max_alignments, alignment_type="global", stream=None, device_id=0,
                 max_device_memory_allocator_caching_size=-1, *args, **kwargs):
        if alignment_type != "global":
            raise RuntimeError("Unknown alignment_type provided. Must be global.")
        if stream is not None and not (hasattr(stream, "stream") or hasattr(stream, "cuda_stream")):
            raise RuntimeError("Type for stream option must be CudaStream")
        self.max_query_length = max_query_length
        self.max_target_length = max_target_length
        self.max_alignments = max_alignments
        # the deprecated factory builds AlignerGlobalHirschbergMyers (cudaaligner/src/aligner.cpp:31-74)
        if max_device_memory_allocator_caching_size < -1:
            raise ValueError("max_device_memory_allocator_caching_size has to be either -1 (=all available GPU memory) or greater or equal than 0.")
        self._aligner = GlobalAligner(max_query_length, max_target_length, max_alignments, "hirschberg_myers", stream=stream, device_id=device_id)
        self.stream = stream

    def add_alignment(self, query, target):
        return self._aligner.add_alignment(query, target)

    def align_all(self):
        self._aligner.align_all()

    def get_alignments(self):
        self._aligner.sync_alignments()
        out = []
        for a in self._aligner.get_alignments():
            out.append(CudaAlignment(a.query, a.target, a.convert_to_cigar(), global_alignment, a.status, a.get_alignment(),
                                     a.format_alignment()))
        return out

    def reset(self):
        self._aligner.reset()
