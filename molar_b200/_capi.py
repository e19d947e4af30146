"""ctypes binding of libmolar_b200.so (include/molar_b200.h).  No torch, no oracle, no CPU fallback:
if the library is missing or no B200 is present the calls raise."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmolar_b200.so")

f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)
u64p = C.POINTER(C.c_uint64)
i64p = C.POINTER(C.c_int64)

MB_OK, MB_ERR_ZERO_MASS, MB_ERR_SIZES, MB_ERR_SVD, MB_ERR_NO_PBC = 0, -1, -2, -3, -4
MB_ERR_BOX, MB_ERR_ARG, MB_ERR_CUDA, MB_ERR_STATE = -5, -6, -7, -8

# every symbol include/molar_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "mb_last_error": (C.c_char_p, []),
    "mb_abi_version": (C.c_int, []),
    "mb_open": (C.c_void_p, [C.c_int]),
    "mb_close": (None, [C.c_void_p]),
    "mb_stream": (C.c_void_p, [C.c_void_p]),
    "mb_synchronize": (C.c_int, [C.c_void_p]),
    "mb_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
    "mb_set_frame": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, f32p]),
    "mb_set_frame_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, f32p]),
    "mb_get_frame": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "mb_set_masses": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "mb_set_frame2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "mb_get_masses": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "mb_search_single": (C.c_int64, [C.c_void_p, C.c_float, u64p, C.c_size_t, C.c_uint8]),
    "mb_search_double": (C.c_int64, [C.c_void_p, C.c_float, u64p, C.c_size_t, u64p, C.c_size_t, C.c_int, C.c_uint8]),
    "mb_search_double_vdw": (C.c_int64, [C.c_void_p, u64p, C.c_size_t, f32p, u64p, C.c_size_t, f32p, C.c_int, C.c_uint8]),
    "mb_search_within": (C.c_int64, [C.c_void_p, C.c_float, u64p, C.c_size_t, u64p, C.c_size_t, C.c_int, C.c_uint8,
                                     f32p, f32p]),
    "mb_count_single": (C.c_int64, [C.c_void_p, C.c_float, u64p, C.c_size_t, C.c_uint8]),
    "mb_fill_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_fill_pairs_u32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_fill_ids": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mb_pairs_device": (C.c_void_p, [C.c_void_p, i64p]),
    "mb_pairs_checksum": (C.c_int, [C.c_void_p, u64p]),
    "mb_last_grid_dims": (C.c_int, [C.c_void_p, u64p]),
    "mb_center_of_mass": (C.c_int, [C.c_void_p, u64p, C.c_size_t, f64p]),
    "mb_gyration": (C.c_int, [C.c_void_p, u64p, C.c_size_t, f64p]),
    "mb_rmsd": (C.c_int, [C.c_void_p, u64p, C.c_size_t, u64p, C.c_size_t, C.c_int, C.c_int, f64p]),
    "mb_fit_transform": (C.c_int, [C.c_void_p, u64p, C.c_size_t, u64p, C.c_size_t, C.c_int, C.c_int, f64p, f64p]),
    "mb_apply_transform": (C.c_int, [C.c_void_p, u64p, C.c_size_t, f64p, f64p]),
    "mb_connectivity": (C.c_int64, [C.c_void_p, C.c_size_t, u64p]),
    "mb_search_connectivity": (C.c_int64, [C.c_void_p, C.c_float, u64p, C.c_size_t, C.c_uint8, u64p]),
    "mb_fill_connectivity": (C.c_int, [C.c_void_p, u64p]),
    "mb_connectivity_checksum": (C.c_int, [C.c_void_p, u64p]),
    "mb_unwrap_connectivity": (C.c_int64, [C.c_void_p, C.c_float, u64p, C.c_size_t, C.c_uint8, i64p]),
    "mb_reduce_many": (C.c_int, [C.c_void_p, u64p, u64p, C.c_size_t, C.c_int, f64p, C.POINTER(C.c_int)]),
    "mb_center_of_geometry": (C.c_int, [C.c_void_p, u64p, C.c_size_t, f64p]),
    "mb_center_pbc": (C.c_int, [C.c_void_p, u64p, C.c_size_t, C.c_int, C.c_uint8, f64p]),
    "mb_gyration_pbc": (C.c_int, [C.c_void_p, u64p, C.c_size_t, f64p]),
    "mb_inertia": (C.c_int, [C.c_void_p, u64p, C.c_size_t, C.c_int, f64p, f64p]),
    "mb_principal_transform": (C.c_int, [C.c_void_p, u64p, C.c_size_t, C.c_int, f64p, f64p]),
    "mb_batch_synth": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_size_t, C.c_size_t, f32p, C.c_int]),
    "mb_batch_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, f32p]),
    "mb_batch_synth_masses": (C.c_int, [C.c_void_p, C.c_uint64, C.c_size_t]),
    "mb_batch_select": (C.c_int, [C.c_void_p, C.c_size_t]),
    "mb_traj_probe": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "mb_batch_load_traj": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, f32p, f32p]),
    "mb_batch_download": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
    "mb_batch_search": (C.c_int, [C.c_void_p, C.c_float, C.c_uint8, C.c_size_t, C.c_size_t, C.c_int, i64p, u64p]),
    "mb_stream_search": (C.c_int, [C.c_void_p, C.c_float, C.c_uint8, C.c_void_p, C.c_size_t, C.c_size_t, f32p, C.c_int, i64p]),
    "mb_stream_fit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, f64p]),
    "mb_stream_pipeline": (C.c_int, [C.c_void_p, C.c_float, C.c_uint8, C.c_void_p, C.c_size_t, C.c_size_t, f32p, f64p]),
    "mb_batch_fit": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, f64p]),
    "mb_batch_pipeline": (C.c_int, [C.c_void_p, C.c_float, C.c_uint8, C.c_size_t, C.c_size_t, f64p]),
    "mb_batch_scalars_device": (C.c_void_p, [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "mb_box_describe": (C.c_int, [f32p, C.POINTER(C.c_int), f32p, f32p, C.POINTER(C.c_uint32)]),
    "mb_plan_describe": (C.c_int, [f32p, C.c_float, C.c_uint8, C.c_size_t, C.c_int, C.POINTER(C.c_int), C.c_void_p, f32p]),
    "mb_comm_unique_id": (C.c_int, [C.c_void_p]),
    "mb_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "mb_comm_init_all": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "mb_comm_destroy": (None, [C.c_void_p]),
    "mb_comm_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mb_gather_scalars": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
    "mb_comm_max": (C.c_int, [C.c_void_p, f64p, C.c_size_t]),
    "mb_comm_barrier": (C.c_int, [C.c_void_p]),
    "mb_timer_record": (C.c_int, [C.c_void_p, C.c_int]),
    "mb_timer_elapsed_ms": (C.c_int, [C.c_void_p, C.c_int, C.c_int, f64p]),
    "mb_host_alloc": (C.c_void_p, [C.c_size_t]),
    "mb_host_free": (None, [C.c_void_p]),
    "mb_launch_count": (C.c_uint64, [C.c_void_p]),
    "mb_get_stat": (C.c_int, [C.c_void_p, C.c_char_p, f64p]),
}

_lib = None


class MolarB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def load():
    """dlopen the in-tree library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        path = os.environ.get("MOLAR_B200_PLUGIN", LIB_PATH)
        if not os.path.exists(path):
            raise MolarB200Error(MB_ERR_STATE, f"{path} not found: build it with `python -m molar_b200.build`")
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    return load().mb_last_error().decode("utf-8", "replace")


def check(rc):
    if rc < 0:
        raise MolarB200Error(int(rc), last_error())
    return rc
