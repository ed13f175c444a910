"""ctypes binding of libmre_b200.so (include/mre_b200.h).  The library is the product; this file only
declares prototypes, builds the .so in-tree when asked, and turns error codes into exceptions.

There is no fallback: if the shared library is missing, import of the compute path fails loudly.
"""
import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
BUILD_DIR = os.path.join(_HERE, "build")
LIB_PATH = os.environ.get("MRE_B200_LIB") or os.path.join(BUILD_DIR, "libmre_b200.so")   # override: A/B timing of two builds
INCLUDE = os.path.join(os.path.dirname(_HERE), "include", "mre_b200.h")

SOURCES = ["index.cpp", "tma_host.cpp", "abi.cu", "transe_rank.cu", "metrics.cu", "sampler.cu", "train_step.cu", "bilinear_rank.cu", "zsl_rank.cu", "peer.cu", "project.cu", "rotate_rank.cu", "index_build.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-O2,-pthread", "-shared"]

# constants mirrored from include/mre_b200.h
OK = 0
TRANSE, DISTMULT, COMPLEX, ROTATE = 0, 1, 2, 3
FILTER_NONE, FILTER_INDEX, FILTER_CSR = 0, 1, 2
RANK_STRICT, RANK_TIES_HALF, RANK_PESSIMISTIC = 0, 1, 2
TOTAL_ENTITY, TOTAL_RELATION, TOTAL_TRAIN, TOTAL_VALID, TOTAL_TEST, TOTAL_TRIPLE = range(6)
SPLIT_TRAIN, SPLIT_VALID, SPLIT_TEST = 0, 1, 2
LOSS_MARGIN, LOSS_SIGMOID, LOSS_SOFTPLUS = 0, 1, 2
PEER_HANDLE_BYTES = 64
PROJECT_TRANSH, PROJECT_TRANSD = 0, 1


class MreError(RuntimeError):
    pass


class ZslModel(C.Structure):
    """struct mre_zsl_model"""
    _fields_ = [("D", C.c_int64), ("symbol_emb", C.c_void_p),
                ("gcn_w", C.c_void_p), ("gcn_b", C.c_void_p), ("fc1_w", C.c_void_p), ("fc1_b", C.c_void_p),
                ("fc2_w", C.c_void_p), ("fc2_b", C.c_void_p), ("reshape_w", C.c_void_p), ("reshape_b", C.c_void_p),
                ("proj1_w", C.c_void_p), ("proj1_b", C.c_void_p), ("proj2_w", C.c_void_p), ("proj2_b", C.c_void_p),
                ("ln_g", C.c_void_p), ("ln_b", C.c_void_p), ("ln_eps", C.c_float)]


class RankJob(C.Structure):
    """struct mre_rank_job"""
    _fields_ = [
        ("ent", C.c_void_p), ("rel", C.c_void_p), ("ent_im", C.c_void_p), ("rel_im", C.c_void_p),
        ("E", C.c_int64), ("R", C.c_int64), ("D", C.c_int64),
        ("scorer", C.c_int32), ("p_norm", C.c_int32), ("normalize", C.c_int32), ("filter", C.c_int32),
        ("q_h", C.c_void_p), ("q_t", C.c_void_p), ("q_r", C.c_void_p), ("q_side", C.c_void_p),
        ("side", C.c_int32), ("n_groups", C.c_int32), ("Q", C.c_int64),
        ("group_qptr", C.c_void_p), ("group_cptr", C.c_void_p), ("cand_idx", C.c_void_p),
        ("filt_ptr", C.c_void_p), ("filt_idx", C.c_void_p), ("filt_nnz", C.c_int64),
        ("counts", C.c_void_p), ("rotate_phase_div", C.c_float),
    ]


def sources_newer_than_lib():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [INCLUDE]
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=False, force=False):
    """Compile csrc/ into build/libmre_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    if os.environ.get("MRE_B200_LIB") or (not force and not sources_newer_than_lib()):
        return LIB_PATH
    os.makedirs(BUILD_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise MreError("nvcc failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


_lib = None


def lib():
    """The loaded library with prototypes set.  Never builds implicitly on a box without sources newer than the
    .so; raises if the .so is absent (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MreError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(the product has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, u64, u32, f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_uint32, C.c_float
    P = C.POINTER
    L.mre_last_error.restype = C.c_char_p
    L.mre_abi_version.restype = i32
    L.mre_device_ok.argtypes = [i32]
    L.mre_index_create.argtypes = [i64, i64] + [vp, vp, vp, i64] * 3 + [P(vp)]
    L.mre_index_create_from_dir.argtypes = [C.c_char_p, P(vp)]
    L.mre_index_create_device.argtypes = [i32, i64, i64] + [vp, vp, vp, i64] * 3 + [P(vp), P(C.c_double)]
    L.mre_index_create_from_dir_device.argtypes = [C.c_char_p, i32, P(vp), P(C.c_double)]
    L.mre_index_device_column.argtypes = [vp, i32, vp]
    L.mre_index_device_column.restype = i64
    L.mre_index_destroy.argtypes = [vp]
    L.mre_index_destroy.restype = None
    L.mre_index_to_device.argtypes = [vp, i32]
    L.mre_index_total.argtypes = [vp, i32]
    L.mre_index_total.restype = i64
    L.mre_index_get_split.argtypes = [vp, i32, vp, vp, vp]
    L.mre_index_get_means.argtypes = [vp, vp, vp]
    L.mre_index_find.argtypes = [vp, i64, i64, i64]
    L.mre_index_load_type_constrain.argtypes = [vp, C.c_char_p]
    L.mre_index_set_type_constrain.argtypes = [vp, vp, vp, vp, vp]
    L.mre_index_type_total.argtypes = [vp, i32]
    L.mre_index_type_total.restype = i64
    L.mre_index_get_type_constrain.argtypes = [vp, i32, vp, vp]
    L.mre_ctx_create.argtypes = [i32, P(vp)]
    L.mre_ctx_destroy.argtypes = [vp]
    L.mre_ctx_destroy.restype = None
    L.mre_ctx_sm_count.argtypes = [vp]
    L.mre_ctx_launch_count.argtypes = [vp]
    L.mre_ctx_launch_count.restype = i64
    L.mre_ctx_option.argtypes = [vp, C.c_char_p, i64]
    L.mre_ctx_stat.argtypes = [vp, C.c_char_p, P(i64)]
    L.mre_rank.argtypes = [vp, vp, P(RankJob), vp]
    L.mre_rank_host.argtypes = [vp, vp, P(RankJob), vp]
    L.mre_predict.argtypes = [vp, P(RankJob), i64, vp, vp]
    L.mre_bilinear_scores.argtypes = [vp, P(RankJob), vp, vp]
    L.mre_metrics.argtypes = [vp, vp, vp, i32, i64, i32, i32, vp, vp, vp, i64, vp]
    samp = [vp, vp, u64, u64, u32, i64, i64, i32, i32, vp, vp, vp, vp, vp]
    L.mre_sample.argtypes = samp
    L.mre_sample_host.argtypes = samp
    L.mre_sample_subgraph.argtypes = [vp, vp, u64, u64, u32, vp, vp, vp, i64, vp, i64, vp, i64, i64, i32, i32, vp, vp, vp, vp]
    L.mre_corrupt_typed.argtypes = [vp, vp, u64, u64, u32, vp, vp, i64, vp, vp]
    L.mre_zsl_entity_features.argtypes = [vp, P(ZslModel), vp, vp, vp, i64, i32, vp, vp, vp]
    L.mre_zsl_rank.argtypes = [vp, P(ZslModel), vp, vp, i64, vp, vp, vp, vp, i64, i64, vp, i64, i32, vp, vp, vp]
    L.mre_transe_margin_step.argtypes = [vp, vp, vp, i64, i64, i64, vp, vp, vp, i64, i64, f32, i32, i32, vp, vp, vp, vp, vp]
    L.mre_sgd_update.argtypes = [vp, vp, vp, i64, f32, vp]
    L.mre_score_triples.argtypes = [vp, i32, vp, vp, vp, vp, i64, vp, vp, vp, i64, i32, i32, vp, vp]
    L.mre_transe_backward.argtypes = [vp, vp, vp, i64, vp, vp, vp, i64, i32, i32, vp, vp, vp, vp, vp]
    L.mre_bilinear_backward.argtypes = [vp, i32, vp, vp, vp, vp, i64, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp]
    L.mre_ns_loss.argtypes = [vp, i32, vp, i64, i64, f32, i32, f32, vp, vp, vp]
    L.mre_ns_train_step.argtypes = [vp, i32, vp, vp, vp, vp, i64, vp, vp, vp, i64, i64, i32, f32, i32, f32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    L.mre_relation_project.argtypes = [vp, i32, vp, vp, vp, vp, i64, i64, i64, i32, vp, vp]
    L.mre_peer_group_create.argtypes = [vp, i32, i32, i64, P(vp), C.c_char_p]
    L.mre_peer_group_connect.argtypes = [vp, C.c_char_p]
    L.mre_peer_group_connect_local.argtypes = [vp, P(vp)]
    L.mre_peer_group_destroy.argtypes = [vp]
    L.mre_peer_group_destroy.restype = None
    L.mre_peer_weights.argtypes = [vp]
    L.mre_peer_weights.restype = vp
    L.mre_peer_grads.argtypes = [vp]
    L.mre_peer_grads.restype = vp
    L.mre_dp_sgd_step.argtypes = [vp, vp, f32, i32, vp]
    L.mre_peer_allreduce_i64.argtypes = [vp, vp, vp, i32, vp]
    L.mre_peer_group_error.argtypes = [vp]
    L.mre_probe_fp32_peak.argtypes = [vp, P(C.c_double)]
    L.mre_probe_mufu_peak.argtypes = [vp, P(C.c_double)]
    L.mre_probe_tf32_peak.argtypes = [vp, P(C.c_double)]
    L.mre_probe_bf16_peak.argtypes = [vp, P(C.c_double)]
    L.mre_ctx_timing.argtypes = [vp, i32]
    L.mre_ctx_timing_read.argtypes = [vp, P(C.c_double), P(i64)]
    _lib = L
    return L


def check(rc):
    if rc != OK:
        raise MreError(f"libmre_b200 error {rc}: {lib().mre_last_error().decode(errors='replace')}")


def exported_symbols_in_header():
    """Names of the functions include/mre_b200.h declares (used by the CPU symbol-export test)."""
    import re
    text = open(INCLUDE).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mre_[a-z0-9_]+)\s*\(", text)))
