"""ctypes loader for libdentist_b200.so (the C ABI declared in include/dentist_b200.h).

Fails loudly when the library is missing: there is no CPU or PyTorch fallback anywhere in this
package.  PyTorch is only used by bench.py / tests for device selection and torch.distributed.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdentist_b200.so")

EXPORTS = [
    "dn_init", "dn_shutdown", "dn_last_error", "dn_version", "dn_launch_count",
    "dn_align_params_default", "dn_las_free", "dn_block_upload", "dn_block_free", "dn_block_bases",
    "dn_align_blocks", "dn_align_host", "dn_las_write", "dn_las_read", "dn_dalign", "dn_damap",
    "dn_las_filter_error", "dn_las_filter_pileup", "dn_compute_qvs", "dn_compute_qvs_v", "dn_dust_block", "dn_block_mask_dust", "dn_block_index", "dn_mask_coverage", "dn_propagate_mask", "dn_dbdust", "dn_las_chain_mapper", "dn_las_chain", "dn_las_merge_device", "dn_block_crop", "dn_consensus_db", "dn_collect_filter", "dn_las_force_flat", "dn_reference_read_candidates", "dn_free", "dn_consensus", "dn_seq_free",
    "dn_process_pileups", "dn_pileup_params_default", "dn_insertion_free", "dn_pile_status_string", "dn_block_add_mask",
    "dn_comm_get_id", "dn_comm_init", "dn_comm_shutdown", "dn_comm_rank", "dn_comm_size", "dn_align_blocks_gather", "dn_align_host_gather",
    "dn_comm_allgatherv", "dn_las_keep_best_chains", "dn_las_transpose", "dn_comm_shared_segment_bytes",
    "dn_compute_qvs_db", "dn_read_qvs_db", "dn_las_bridge",
]


class DnError(RuntimeError):
    """Raised for any non-zero return of the C ABI (the D shim raises DazzlerCommandException)."""


class LasRecord(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("tlen", "diffs", "abpos", "bbpos", "aepos", "bepos")] + \
               [("flags", C.c_uint32), ("aread", C.c_int32), ("bread", C.c_int32), ("pad_", C.c_int32)]


REC_DTYPE = np.dtype([("tlen", "<i4"), ("diffs", "<i4"), ("abpos", "<i4"), ("bbpos", "<i4"),
                      ("aepos", "<i4"), ("bepos", "<i4"), ("flags", "<u4"), ("aread", "<i4"),
                      ("bread", "<i4"), ("pad", "<i4")])


class BlockDesc(C.Structure):
    _fields_ = [("nreads", C.c_int32), ("format", C.c_int32), ("rlen", C.c_void_p), ("boff", C.c_void_p),
                ("data", C.c_void_p), ("data_bytes", C.c_int64), ("mask_anno", C.c_void_p), ("mask_data", C.c_void_p),
                ("group", C.c_void_p)]


class AlignParams(C.Structure):
    _fields_ = [("k", C.c_int32), ("w", C.c_int32), ("h", C.c_int32), ("t", C.c_int32), ("tspace", C.c_int32),
                ("minlen", C.c_int32), ("e", C.c_double), ("identity", C.c_int32), ("self_block", C.c_int32),
                ("rounds", C.c_int32), ("xdrop", C.c_int32), ("wmax", C.c_int32), ("poolmul", C.c_int32),
                ("join_mode", C.c_int32)]


class AlignStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("tuples_a", "tuples_b", "hits", "seeds", "extensions", "las",
                                         "aligned_bases", "trace_points", "algo_bytes_seed", "algo_bytes_extend")] + \
               [("ms_seed", C.c_float), ("ms_extend", C.c_float), ("ms_total", C.c_float), ("launches", C.c_uint64)]


class LasBuf(C.Structure):
    _fields_ = [("nrec", C.c_int64), ("rec", C.POINTER(LasRecord)), ("toff", C.POINTER(C.c_int64)),
                ("ntrace", C.c_int64), ("trace", C.POINTER(C.c_uint16)), ("tspace", C.c_int32), ("stats", AlignStats)]


class PileupDesc(C.Structure):
    _fields_ = [("nreads", C.c_int32), ("rlen", C.c_void_p), ("bases", C.c_void_p), ("allowed", C.c_void_p),
                ("nflanks", C.c_int32), ("flank_read", C.c_void_p), ("mask_anno", C.c_void_p), ("mask_data", C.c_void_p)]


class PileupParams(C.Structure):
    _fields_ = [("max_alignment_error", C.c_double), ("min_anchor_length", C.c_int32), ("tspace", C.c_int32),
                ("proper_alignment_allowance", C.c_int32), ("bad_fraction", C.c_double), ("min_qv_coverage", C.c_int32),
                ("dust", C.c_int32), ("max_indel", C.c_int32), ("max_chain_gap", C.c_int32), ("max_rel_overlap", C.c_double),
                ("min_rel_score", C.c_double), ("min_score", C.c_int32), ("k", C.c_int32), ("flank_k", C.c_int32), ("bridge", C.c_int32)]


class InsertionOut(C.Structure):
    _fields_ = [("status", C.c_int32), ("reference_read", C.c_int32), ("ntries", C.c_int32), ("cons_len", C.c_int64),
                ("consensus", C.POINTER(C.c_uint8)), ("flank_las", LasBuf)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run ./build.sh (or __graft_entry__.build()); dentist_b200 has no "
                              "CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.dn_last_error.restype = C.c_char_p
        L.dn_version.restype = C.c_char_p
        L.dn_launch_count.restype = C.c_uint64
        L.dn_init.argtypes = [C.c_int, C.c_char_p]
        L.dn_align_params_default.argtypes = [C.POINTER(AlignParams)]
        L.dn_las_free.argtypes = [C.POINTER(LasBuf)]
        L.dn_block_upload.argtypes = [C.POINTER(BlockDesc), C.POINTER(C.c_void_p)]
        L.dn_block_crop.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.dn_block_free.argtypes = [C.c_void_p]
        L.dn_block_bases.argtypes = [C.c_void_p]
        L.dn_block_bases.restype = C.c_int64
        L.dn_align_blocks.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(AlignParams), C.POINTER(LasBuf)]
        L.dn_align_host.argtypes = [C.POINTER(BlockDesc), C.POINTER(BlockDesc), C.POINTER(AlignParams), C.POINTER(LasBuf)]
        L.dn_las_write.argtypes = [C.c_char_p, C.POINTER(LasBuf)]
        L.dn_las_read.argtypes = [C.c_char_p, C.POINTER(LasBuf)]
        L.dn_dalign.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_char_p]
        L.dn_damap.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_char_p]
        L.dn_las_filter_error.argtypes = [C.POINTER(LasBuf), C.c_double]
        L.dn_las_filter_pileup.argtypes = [C.POINTER(LasBuf), C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32]
        L.dn_compute_qvs.argtypes = [C.c_void_p, C.c_int32, C.POINTER(LasBuf), C.c_int32, C.c_void_p, C.c_void_p]
        L.dn_compute_qvs_v.argtypes = [C.c_void_p, C.c_int32, C.POINTER(LasBuf), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.dn_las_merge_device.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int64,
                                          C.c_int64, C.POINTER(LasBuf)]
        L.dn_las_chain.argtypes = [C.POINTER(LasBuf), C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32]
        L.dn_las_chain_mapper.argtypes = [C.POINTER(LasBuf), C.c_int32, C.c_int32, C.c_int32]
        L.dn_las_keep_best_chains.argtypes = [C.POINTER(LasBuf), C.c_int32, C.c_double]
        L.dn_las_transpose.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(LasBuf), C.POINTER(LasBuf)]
        L.dn_consensus_db.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32, C.POINTER(C.c_char_p), C.c_int, C.c_char_p, C.c_size_t]
        L.dn_collect_filter.argtypes = [C.POINTER(LasBuf), C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_double,
                                        C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.dn_las_force_flat.argtypes = [C.POINTER(LasBuf)]
        L.dn_reference_read_candidates.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_void_p]
        L.dn_dbdust.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int]
        L.dn_las_bridge.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(LasBuf), C.c_int32, C.c_void_p]
        L.dn_compute_qvs_db.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32]
        L.dn_read_qvs_db.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.dn_dust_block.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_int32, C.c_void_p, C.c_void_p]
        L.dn_mask_coverage.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.dn_propagate_mask.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.dn_block_index.argtypes = [C.c_void_p, C.c_int32]
        L.dn_block_mask_dust.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_int32, C.c_void_p]
        L.dn_free.argtypes = [C.c_void_p]
        L.dn_consensus.argtypes = [C.c_void_p, C.POINTER(LasBuf), C.c_void_p, C.c_int32, C.c_void_p]
        L.dn_seq_free.argtypes = [C.c_void_p]
        L.dn_pileup_params_default.argtypes = [C.POINTER(PileupParams)]
        L.dn_process_pileups.argtypes = [C.c_void_p, C.POINTER(PileupDesc), C.c_int32, C.POINTER(PileupParams), C.POINTER(InsertionOut)]
        L.dn_insertion_free.argtypes = [C.POINTER(InsertionOut), C.c_int32]
        L.dn_pile_status_string.argtypes = [C.c_int32]
        L.dn_pile_status_string.restype = C.c_char_p
        L.dn_block_add_mask.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.dn_comm_get_id.argtypes = [C.c_void_p]
        L.dn_comm_shared_segment_bytes.restype = C.c_int64
        L.dn_comm_init.argtypes = [C.c_int32, C.c_int32, C.c_void_p]
        L.dn_align_blocks_gather.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(AlignParams), C.c_int64, C.c_int32, C.POINTER(LasBuf)]
        L.dn_align_host_gather.argtypes = [C.POINTER(BlockDesc), C.POINTER(BlockDesc), C.POINTER(AlignParams), C.c_int64, C.c_int32, C.POINTER(LasBuf)]
        L.dn_comm_allgatherv.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.c_void_p]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise DnError("dentist_b200 error %d: %s" % (rc, lib().dn_last_error().decode("utf-8", "replace")))
