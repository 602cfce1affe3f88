"""ctypes loader for libhlacount.so. Fails loudly when the library is absent (no fallback of any kind)."""
import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libhlacount.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "hlacount.h")
_LIB = None


class HlacError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("hlacount error %d: %s" % (code, msg))
        self.code = code


def build(jobs=8):
    """Compile the library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-C", _HERE, "-j%d" % jobs, "-s"])
    return SO_PATH


def declared_symbols():
    """Every function include/hlacount.h declares."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hlac_[a-z0-9_]+)\s*\(", txt)))


class Batch(C.Structure):
    _fields_ = [("seq", C.c_void_p), ("len", C.c_void_p), ("cell", C.c_void_p), ("umi", C.c_void_p),
                ("flags", C.c_void_p), ("n", C.c_uint64), ("words_per_read", C.c_uint32), ("ordinal", C.c_void_p)]


class PackedBatch(C.Structure):
    _fields_ = [("rec", C.c_void_p), ("n", C.c_uint64), ("words_per_read", C.c_uint32)]


class WireBatch(C.Structure):
    _fields_ = [("rec", C.c_void_p), ("n", C.c_uint64), ("read_len", C.c_uint32), ("cell_bits", C.c_uint32), ("umi_bits", C.c_uint32),
                ("words_per_record", C.c_uint32)]


class Metrics(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("num_reads", "num_non_primary", "num_not_cell_bc", "num_cds_align",
                                          "num_gen_align", "num_not_aligned")]


class Coo(C.Structure):
    _fields_ = [("row", C.POINTER(C.c_uint32)), ("col", C.POINTER(C.c_uint32)), ("val", C.POINTER(C.c_uint32)),
                ("n", C.c_uint64), ("nrows", C.c_uint32), ("borrowed", C.c_uint32)]


class IndexInfo(C.Structure):
    _fields_ = [("n_nodes", C.c_uint32), ("n_tx", C.c_uint32), ("n_colours", C.c_uint32), ("lmer", C.c_uint32),
                ("n_bases", C.c_uint64), ("n_kmers", C.c_uint64), ("dict_slots", C.c_uint64),
                ("bitmap_bytes", C.c_uint64)]


class ClassTables(C.Structure):
    _fields_ = [("nitems", C.c_uint32), ("nreads", C.c_uint32),
                ("n_read_classes", C.c_uint64), ("reads_off", C.POINTER(C.c_uint64)),
                ("reads_tx", C.POINTER(C.c_uint32)), ("reads_cnt", C.POINTER(C.c_uint32)),
                ("n_umi_classes", C.c_uint64), ("umi_off", C.POINTER(C.c_uint64)),
                ("umi_tx", C.POINTER(C.c_uint32)), ("umi_cnt", C.POINTER(C.c_uint32)),
                ("reads_explained", C.POINTER(C.c_uint64))]


class EqClassDbC(C.Structure):
    _fields_ = [("n_barcodes", C.c_uint32), ("barcode_off", C.POINTER(C.c_uint64)), ("barcode_bytes", C.POINTER(C.c_uint8)),
                ("n_umis", C.c_uint32), ("umi_off", C.POINTER(C.c_uint64)), ("umi_bytes", C.POINTER(C.c_uint8)),
                ("n_classes", C.c_uint32), ("class_off", C.POINTER(C.c_uint64)), ("class_tx", C.POINTER(C.c_uint32)),
                ("n_counts", C.c_uint64), ("bc", C.POINTER(C.c_uint32)), ("umi", C.POINTER(C.c_uint32)),
                ("cls", C.POINTER(C.c_uint32))]


class Paths(C.Structure):
    _fields_ = [("n_paths", C.c_uint64), ("hash_a", C.POINTER(C.c_uint64)), ("hash_b", C.POINTER(C.c_uint64)),
                ("first_ord", C.POINTER(C.c_uint64)), ("count", C.POINTER(C.c_uint32)), ("off", C.POINTER(C.c_uint64)),
                ("colours", C.POINTER(C.c_uint32))]


class Calls(C.Structure):
    _fields_ = [("n_called", C.c_uint32), ("called", C.POINTER(C.c_uint32)),
                ("n_weights", C.c_uint32), ("w_tx", C.POINTER(C.c_uint32)), ("w_em", C.POINTER(C.c_double)),
                ("w_frac", C.POINTER(C.c_double)),
                ("n_pairs", C.c_uint32), ("p_a", C.POINTER(C.c_uint32)), ("p_b", C.POINTER(C.c_uint32)),
                ("p_frac", C.POINTER(C.c_double))]


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise HlacError(-4, "libhlacount.so is not built (%s); run `python -c 'import __graft_entry__ as g; "
                                "g.build()'` — there is no fallback path" % SO_PATH)
        L = C.CDLL(SO_PATH)
        L.hlac_last_error.restype = C.c_char_p
        L.hlac_index_tx_name.restype = C.c_char_p
        L.hlac_index_tx_name.argtypes = [C.c_void_p, C.c_uint32]
        L.hlac_count_row_name.restype = C.c_char_p
        L.hlac_count_row_name.argtypes = [C.c_void_p, C.c_uint32]
        L.hlac_count_nrows.restype = C.c_uint32
        L.hlac_count_nrows.argtypes = [C.c_void_p]
        L.hlac_wire_words.restype = C.c_uint32
        L.hlac_wire_words.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        raise HlacError(rc, lib().hlac_last_error().decode(errors="replace"))
