"""Loads libmcbb_b200.so (the CUDA C-ABI library) and declares its prototypes.

There is no CPU fallback: if the shared library is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

from . import _abi as abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmcbb_b200.so")

# every symbol include/mcbb_b200.h declares
EXPORTS = [
    "mcbb_version", "mcbb_init", "mcbb_finalize", "mcbb_last_error", "mcbb_device_info",
    "mcbb_kernel_launch_count", "mcbb_tsave_array", "mcbb_solve_features", "mcbb_solve_features_dev",
    "mcbb_solve_features_saved", "mcbb_solve_features_saved_dev", "mcbb_set_solver_parameter_var",
    "mcbb_knn_dist_features", "mcbb_knn_dist_features_dev",
    "mcbb_distance_matrix", "mcbb_distance_matrix_dev", "mcbb_dbscan_dense", "mcbb_dbscan_dense_dev",
    "mcbb_dbscan_features", "mcbb_dbscan_features_dev", "mcbb_dbscan_rows_count_dev",
    "mcbb_dbscan_rows_union_dev", "mcbb_dbscan_merge_parents_dev", "mcbb_dbscan_rows_border_dev",
    "mcbb_dbscan_labels_dev", "mcbb_gather_columns_dev", "mcbb_measure_fp64_fma_peak", "mcbb_knn_dist_dense", "mcbb_knn_dist_dense_dev",
    "mcbb_distance_sparse_count", "mcbb_distance_sparse_count_dev", "mcbb_distance_sparse_fill", "mcbb_distance_sparse_fill_dev",
    "mcbb_dbscan_sparse", "mcbb_dbscan_sparse_dev", "mcbb_dbscan_rows_count_lists_dev", "mcbb_dbscan_rows_union_more_dev", "mcbb_distance_matrix_rows", "mcbb_distance_matrix_rows_dev",
    "mcbb_test_hook_set", "mcbb_pair_feature_counter",
    "mcbb_multi_init", "mcbb_multi_finalize", "mcbb_multi_info", "mcbb_solve_features_multi", "mcbb_dbscan_features_multi",
    "mcbb_solve_cluster_multi",
    "mcbb_order_stats", "mcbb_order_stats_dev", "mcbb_hist_rows", "mcbb_hist_rows_dev",
]

_lib = None


class MCBBError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MCBBError("libmcbb_b200.so is not built (run ./build.sh or __graft_entry__.build()); "
                        "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    dp, ip, lp, vp = abi.c_double_p, abi.c_int32_p, abi.c_int64_p, C.c_void_p
    SD, SO, FO = C.POINTER(abi.SystemDesc), C.POINTER(abi.SolverOpts), C.POINTER(abi.FeatOpts)
    L.mcbb_version.restype = C.c_char_p
    L.mcbb_init.argtypes = [C.c_int]
    L.mcbb_last_error.argtypes = [C.c_char_p, C.c_size_t]
    L.mcbb_last_error.restype = C.c_size_t
    L.mcbb_device_info.argtypes = [ip, ip, ip, lp]
    L.mcbb_kernel_launch_count.restype = C.c_int64
    L.mcbb_tsave_array.argtypes = [SO, C.c_int32, C.c_double, dp, ip]
    L.mcbb_solve_features.argtypes = [SD, SO, FO, C.c_int64, dp, dp, dp, C.c_int32, dp, dp, dp, ip, lp]
    # device entry points take raw addresses
    L.mcbb_solve_features_dev.argtypes = [SD, SO, FO, C.c_int64, vp, vp, vp, C.c_int32, vp, vp, vp, vp, vp, vp]
    L.mcbb_solve_features_saved.argtypes = [SD, SO, FO, C.c_int64, dp, dp, dp, C.c_int32, dp, dp, dp, ip, lp,
                                            C.c_int32, dp]
    L.mcbb_set_solver_parameter_var.argtypes = [C.c_int32, dp, C.c_int64]
    L.mcbb_solve_features_saved_dev.argtypes = [SD, SO, FO, C.c_int64, vp, vp, vp, C.c_int32, vp, vp, vp, vp, vp,
                                                C.c_int32, vp, vp]
    L.mcbb_distance_matrix.argtypes = [dp, C.c_int64, C.c_int32, ip, dp, dp]
    L.mcbb_distance_matrix_dev.argtypes = [vp, C.c_int64, C.c_int32, ip, dp, vp, vp]
    L.mcbb_dbscan_dense.argtypes = [dp, C.c_int64, C.c_double, C.c_int32, lp, lp, lp, lp]
    L.mcbb_dbscan_dense_dev.argtypes = [vp, C.c_int64, C.c_double, C.c_int32, vp, vp, vp, lp, vp]
    L.mcbb_dbscan_features.argtypes = [dp, C.c_int64, C.c_int32, ip, dp, C.c_double, C.c_int32, lp, lp, lp, lp]
    L.mcbb_dbscan_features_dev.argtypes = [vp, C.c_int64, C.c_int32, ip, dp, C.c_double, C.c_int32, vp, vp, vp,
                                           lp, vp]
    L.mcbb_dbscan_rows_count_dev.argtypes = [vp, C.c_int64, C.c_int32, ip, dp, C.c_double, C.c_int32, C.c_int64,
                                             C.c_int64, vp, vp]
    L.mcbb_dbscan_rows_union_dev.argtypes = [vp, C.c_int64, C.c_int32, ip, dp, C.c_double, C.c_int64, C.c_int64,
                                             vp, vp]
    L.mcbb_dbscan_merge_parents_dev.argtypes = [vp, vp, C.c_int64, vp]
    L.mcbb_dbscan_rows_border_dev.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64, vp, C.c_int64, vp, C.c_int32,
                                              ip, dp, C.c_double, vp, vp]
    L.mcbb_dbscan_labels_dev.argtypes = [vp, C.c_int64, vp, vp, vp, lp, vp]
    L.mcbb_measure_fp64_fma_peak.argtypes = [dp, dp]
    L.mcbb_knn_dist_dense.argtypes = [dp, C.c_int64, C.c_int32, C.c_int32, dp, dp]
    L.mcbb_knn_dist_dense_dev.argtypes = [vp, C.c_int64, C.c_int32, C.c_int32, vp, vp, vp]
    L.mcbb_knn_dist_features.argtypes = [dp, C.c_int64, C.c_int32, ip, dp, C.c_int32, C.c_int32, C.c_int64, dp, dp]
    L.mcbb_knn_dist_features_dev.argtypes = [vp, C.c_int64, C.c_int32, ip, dp, C.c_int32, C.c_int32, C.c_int64, vp, vp,
                                             vp]
    L.mcbb_order_stats.argtypes = [dp, C.c_int64, C.c_int32, lp, dp, lp]
    L.mcbb_order_stats_dev.argtypes = [vp, C.c_int64, C.c_int32, lp, dp, lp, vp]
    L.mcbb_hist_rows.argtypes = [dp, C.c_int64, C.c_int32, lp, C.c_int32, C.c_double, C.c_double, dp, C.c_int32, C.c_int32, dp]
    L.mcbb_hist_rows_dev.argtypes = [vp, C.c_int64, C.c_int32, vp, C.c_int32, C.c_double, C.c_double, dp, C.c_int32, C.c_int32,
                                     vp, vp]
    L.mcbb_distance_sparse_count.argtypes = [dp, C.c_int64, C.c_int32, ip, dp, C.c_double, ip]
    L.mcbb_distance_sparse_count_dev.argtypes = [vp, C.c_int64, C.c_int32, ip, dp, C.c_double, vp, vp]
    L.mcbb_distance_sparse_fill.argtypes = [dp, C.c_int64, C.c_int32, ip, dp, C.c_double, lp, ip, dp]
    L.mcbb_distance_sparse_fill_dev.argtypes = [vp, C.c_int64, C.c_int32, ip, dp, C.c_double, vp, C.c_int64, vp, vp, vp]
    L.mcbb_distance_matrix_rows.argtypes = [dp, C.c_int64, C.c_int32, ip, dp, C.c_int64, C.c_int64, dp]
    L.mcbb_distance_matrix_rows_dev.argtypes = [vp, C.c_int64, C.c_int32, ip, dp, C.c_int64, C.c_int64, vp, vp]
    L.mcbb_dbscan_rows_count_lists_dev.argtypes = [vp, C.c_int64, C.c_int32, ip, dp, C.c_double, C.c_int32, C.c_int64, C.c_int64,
                                                   vp, vp, vp, vp]
    L.mcbb_dbscan_rows_union_more_dev.argtypes = [vp, C.c_int64, C.c_int32, ip, dp, C.c_double, C.c_int64, C.c_int64, vp, vp]
    L.mcbb_dbscan_sparse.argtypes = [C.c_int64, lp, ip, dp, C.c_double, C.c_double, C.c_int32, lp, lp, lp, lp]
    L.mcbb_dbscan_sparse_dev.argtypes = [C.c_int64, vp, vp, vp, C.c_double, C.c_double, C.c_int32, vp, vp, vp, lp, vp]
    L.mcbb_multi_init.argtypes = [C.c_int32, ip]
    L.mcbb_multi_info.argtypes = [ip, ip]
    L.mcbb_solve_features_multi.argtypes = [SD, SO, FO, C.c_int64, dp, dp, dp, C.c_int32, dp, dp, dp, ip, lp]
    L.mcbb_dbscan_features_multi.argtypes = [dp, C.c_int64, C.c_int32, ip, dp, C.c_double, C.c_int32, lp, lp, lp, lp]
    L.mcbb_solve_cluster_multi.argtypes = [SD, SO, FO, C.c_int64, dp, dp, dp, C.c_int32, dp, C.c_double, C.c_int32, dp, ip,
                                           lp, lp, lp, lp, lp, lp]
    L.mcbb_pair_feature_counter.argtypes = [C.c_int32, lp, vp]
    L.mcbb_test_hook_set.argtypes = [C.c_char_p, C.c_int64]
    L.mcbb_gather_columns_dev.argtypes = [vp, C.c_int64, vp, C.c_int64, C.c_int32, vp, vp]
    _lib = L
    return L


def last_error():
    buf = C.create_string_buffer(2048)
    lib().mcbb_last_error(buf, 2048)
    return buf.value.decode("utf-8", "replace")


def check(status):
    if status != 0:
        raise MCBBError(last_error() or ("mcbb_b200 call failed with status %d" % status))


def kernel_launch_count():
    return int(lib().mcbb_kernel_launch_count())
