"""ctypes binding of libgrmkm.so (include/grmkm.h).  Fails loudly: no library or no GPU -> error."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GRMKM_LIB") or os.path.join(_HERE, "libgrmkm.so")   # GRMKM_LIB: another build of the same library (kernel experiments)
ABI_VERSION = 10

OK = 0
E_INVALID, E_UNSUPPORTED_K, E_NOMEM, E_CUDA, E_IO, E_CAPACITY, E_NO_DEVICE, E_UNSUPPORTED = -1, -2, -3, -4, -5, -6, -7, -8
FASTA, FASTQ = 0, 1
FLAG_COUNTS = 32      # pooled count table (dsk mode): matrix row 0 = abundance of the k-mer over all inputs

# every symbol include/grmkm.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "grmkm_abi_version", "grmkm_device_count", "grmkm_create", "grmkm_destroy", "grmkm_last_error", "grmkm_reset",
    "grmkm_add_genome_bytes", "grmkm_add_genome_device", "grmkm_add_genome_files", "grmkm_add_genomes", "grmkm_set_genome_count",
    "grmkm_build", "grmkm_dims", "grmkm_get_stats", "grmkm_stage_times", "grmkm_copy_kmers_packed",
    "grmkm_copy_kmer_strings", "grmkm_copy_matrix", "grmkm_format_tsv", "grmkm_device_result", "grmkm_host_result",
    "grmkm_synth_fasta_device", "grmkm_build_partial", "grmkm_export_partials", "grmkm_export_partials_peers", "grmkm_merge_partials",
    "grmkm_plan_bucket_bits", "grmkm_set_bucket_bits", "grmkm_set_exchange_rank",
    "grmkm_result_checksum", "grmkm_sum_rows", "grmkm_gram", "grmkm_tsv_pack",
]


class GrmkmError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libgrmkm error {code}: {message}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("k", C.c_uint32), ("min_abundance", C.c_uint32),
                ("keep_singletons", C.c_uint32), ("input_kind", C.c_uint32), ("device", C.c_int32),
                ("bucket_bits", C.c_uint32), ("flags", C.c_uint32), ("stream", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("n_bases", C.c_uint64), ("n_windows", C.c_uint64), ("n_input_bytes", C.c_uint64),
                ("n_records", C.c_uint64), ("n_kmers", C.c_uint64), ("n_distinct", C.c_uint64),
                ("n_words", C.c_uint32), ("n_genomes", C.c_uint32), ("n_buckets", C.c_uint32),
                ("n_launches", C.c_uint32), ("h2d_bytes", C.c_uint64), ("device_bytes", C.c_uint64),
                ("n_splits", C.c_uint64), ("n_region_overflows", C.c_uint64), ("n_units", C.c_uint64),
                ("n_unit_entries", C.c_uint64), ("n_wide", C.c_uint64), ("n_unit_buckets", C.c_uint32),
                ("n_rounds", C.c_uint32), ("n_solid_records", C.c_uint64)]

    def asdict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class Times(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h2d", "parse", "pack", "count", "scatter", "abundance", "aggregate",
                                         "sort", "total", "bounds", "dedupe", "expand")]

    def asdict(self):
        return {n: float(getattr(self, n)) for n, _ in self._fields_}


_lib = None


def load() -> C.CDLL:
    """Load libgrmkm.so; raise if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                          "(there is no CPU fallback for the k-mer matrix path)")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    sig = {
        "grmkm_abi_version": (i32, []),
        "grmkm_device_count": (i32, []),
        "grmkm_create": (i32, [C.POINTER(Config), C.POINTER(vp)]),
        "grmkm_destroy": (None, [vp]),
        "grmkm_last_error": (C.c_char_p, [vp]),
        "grmkm_reset": (i32, [vp]),
        "grmkm_add_genome_bytes": (i32, [vp, u32, vp, u64]),
        "grmkm_add_genome_device": (i32, [vp, u32, vp, u64]),
        "grmkm_add_genome_files": (i32, [vp, u32, C.POINTER(C.c_char_p), i32]),
        "grmkm_set_genome_count": (i32, [vp, u32]),
        "grmkm_build": (i32, [vp]),
        "grmkm_dims": (i32, [vp, C.POINTER(u64), C.POINTER(u32), C.POINTER(u32)]),
        "grmkm_get_stats": (i32, [vp, C.POINTER(Stats)]),
        "grmkm_stage_times": (i32, [vp, C.POINTER(Times)]),
        "grmkm_copy_kmers_packed": (i32, [vp, vp, u64]),
        "grmkm_copy_kmer_strings": (i32, [vp, vp, u64]),
        "grmkm_copy_matrix": (i32, [vp, vp, u64]),
        "grmkm_add_genomes": (i32, [vp, u32, vp, vp, vp, i32]),
        "grmkm_format_tsv": (i32, [vp, C.POINTER(C.c_char_p), vp, u64, C.POINTER(u64)]),
        "grmkm_device_result": (i32, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64)]),
        "grmkm_host_result": (i32, [vp, C.POINTER(vp), C.POINTER(vp)]),
        "grmkm_synth_fasta_device": (i32, [vp, vp, u64, vp, u64]),
        "grmkm_build_partial": (i32, [vp, u32, C.POINTER(u64)]),
        "grmkm_export_partials": (i32, [vp, vp, u64]),
        "grmkm_export_partials_peers": (i32, [vp, u32, vp, vp]),
        "grmkm_merge_partials": (i32, [vp, vp, u32, u32, C.POINTER(u64), C.POINTER(u32), u32]),
        "grmkm_plan_bucket_bits": (i32, [vp, C.POINTER(u32)]),
        "grmkm_set_bucket_bits": (i32, [vp, u32]),
        "grmkm_set_exchange_rank": (i32, [vp, u32]),
        "grmkm_result_checksum": (i32, [vp, C.POINTER(u64)]),
        "grmkm_sum_rows": (i32, [vp, vp, u32, vp, u64]),
        "grmkm_gram": (i32, [vp, vp, u64]),
        "grmkm_tsv_pack": (i32, [vp, vp, u64, u32, u32, u32, u32, vp, vp, u64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    if L.grmkm_abi_version() != ABI_VERSION:
        raise ImportError(f"libgrmkm ABI {L.grmkm_abi_version()} != expected {ABI_VERSION}; rebuild")
    _lib = L
    return L
