"""ctypes binding of libpauxy_b200.so (include/pauxy_b200.h).

The product path has no CPU fallback: if the CUDA library is missing this
module raises at import of the engine, loudly.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'csrc', 'libpauxy_b200.so')

PXB_OK = 0
ABI_VERSION = 10
EXCHANGE_MODES = {'auto': 0, 'cholesky': 1, 'eri': 2}
ERRORS = {-1: 'PXB_ERR_ARG', -2: 'PXB_ERR_CUDA', -3: 'PXB_ERR_STATE', -4: 'PXB_ERR_UNSUPPORTED'}

# enum pxb_field_id
F_WEIGHT, F_UNSCALED_WEIGHT, F_OT, F_HYBRID_ENERGY, F_ELOC, F_DETR, F_LOG_DETR, \
    F_ESTIMATES, F_COUNTERS, F_PARENT_IX, F_XBAR, F_XSHIFTED, F_CMF_CFB, F_OVLP_NEW, \
    F_TOTAL_WEIGHT, F_PAIRS, F_PHASE, F_BP_RDM, F_BP_DENOM, F_THETA_SUM, F_WALKER_ELOC, F_OVLP_DET, \
    F_LOG_SHIFTS, F_COUNT = range(24)
FLAG_FREE_PROJECTION, FLAG_NO_FORCE_BIAS, FLAG_LOCAL_ENERGY_WEIGHT, FLAG_COMPLEX_ONE_BODY = 1, 2, 4, 8
FLAG_COMPLEX_CHOLESKY = 16
MAX_DETS = 8
STEP_ORTHO, STEP_POP, STEP_ENERGY = 1, 2, 4


STAGES = ['greens', 'xgemm', 'field', 'vhs', 'one_body', 'taylor', 'weight', 'exchange', 'energy',
          'qr', 'pop_control', 'accumulate']


class PxbConfig(ctypes.Structure):
    _fields_ = [('nbasis', ctypes.c_int32), ('nup', ctypes.c_int32), ('ndown', ctypes.c_int32),
                ('nchol', ctypes.c_int32), ('nwalkers', ctypes.c_int32),
                ('exp_order', ctypes.c_int32), ('device', ctypes.c_int32),
                ('total_walkers', ctypes.c_int32), ('dt', ctypes.c_double),
                ('exchange_mode', ctypes.c_int32), ('flags', ctypes.c_int32), ('nbp', ctypes.c_int32),
                ('ndets', ctypes.c_int32)]


class PxbError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "%s (%d): %s" % (ERRORS.get(code, 'PXB_ERR'), code, msg))
        self.code = code


_vp = ctypes.c_void_p
_PROTOS = {
    'pxb_abi_version': (ctypes.c_int, []),
    'pxb_launch_count': (ctypes.c_longlong, [_vp]),
    'pxb_stage_exchange': (ctypes.c_int, [_vp, _vp]),
    'pxb_profile': (ctypes.c_int, [_vp, ctypes.c_int]),
    'pxb_exchange_mode': (ctypes.c_int, [_vp]),
    'pxb_vhs_symmetric': (ctypes.c_int, [_vp]),
    'pxb_stage_times': (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_double),
                                       ctypes.POINTER(ctypes.c_longlong), ctypes.c_int, ctypes.c_int]),
    'pxb_create': (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.POINTER(PxbConfig)]),
    'pxb_destroy': (ctypes.c_int, [_vp]),
    'pxb_last_error': (ctypes.c_char_p, [_vp]),
    'pxb_arena_bytes': (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_size_t)]),
    'pxb_bind_arena': (ctypes.c_int, [_vp, _vp, ctypes.c_size_t, _vp]),
    'pxb_field': (ctypes.c_int, [_vp, ctypes.c_int, ctypes.POINTER(ctypes.c_size_t),
                                 ctypes.POINTER(ctypes.c_size_t)]),
    'pxb_set_hamiltonian': (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_double, _vp]),
    'pxb_set_trial_det': (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_double, ctypes.c_double, _vp, _vp, _vp, _vp]),
    'pxb_set_eshift_imag': (ctypes.c_int, [_vp, ctypes.c_double]),
    'pxb_set_phi': (ctypes.c_int, [_vp, _vp, _vp]),
    'pxb_get_phi': (ctypes.c_int, [_vp, _vp, _vp]),
    'pxb_init_walkers': (ctypes.c_int, [_vp, _vp, ctypes.c_double, _vp]),
    'pxb_propagate': (ctypes.c_int, [_vp, _vp, ctypes.c_uint64, ctypes.c_int64, ctypes.c_double,
                                     ctypes.c_int64, _vp]),
    'pxb_orthogonalise': (ctypes.c_int, [_vp, _vp]),
    'pxb_step': (ctypes.c_int, [_vp, _vp, ctypes.c_uint64, ctypes.c_int64, ctypes.c_double,
                                ctypes.c_int64, ctypes.c_double, ctypes.c_int, _vp]),
    'pxb_step_graphs': (ctypes.c_int, [_vp, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong)]),
    'pxb_local_energy': (ctypes.c_int, [_vp, _vp]),
    'pxb_accumulate': (ctypes.c_int, [_vp, ctypes.c_int, _vp]),
    'pxb_zero_estimates': (ctypes.c_int, [_vp, _vp]),
    'pxb_accumulate_theta': (ctypes.c_int, [_vp, _vp]),
    'pxb_log_shift_enable': (ctypes.c_int, [_vp, ctypes.c_int, _vp]),
    'pxb_log_shift_sums': (ctypes.c_int, [_vp, _vp]),
    'pxb_log_shift_update': (ctypes.c_int, [_vp, _vp]),
    'pxb_pop_control_comb': (ctypes.c_int, [_vp, ctypes.c_double, _vp]),
    'pxb_pop_rescale': (ctypes.c_int, [_vp, _vp, ctypes.c_int64, _vp]),
    'pxb_comb_plan': (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_double, _vp]),
    'pxb_payload_doubles': (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_size_t)]),
    'pxb_copy_walkers': (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int, _vp]),
    'pxb_pack_walkers': (ctypes.c_int, [_vp, _vp, ctypes.c_int, _vp, _vp]),
    'pxb_unpack_walkers': (ctypes.c_int, [_vp, _vp, ctypes.c_int, _vp, _vp]),
    'pxb_set_weights': (ctypes.c_int, [_vp, ctypes.c_double, _vp]),
    'pxb_peer_export': (ctypes.c_int, [_vp, _vp, ctypes.POINTER(ctypes.c_uint64)]),
    'pxb_peer_attach': (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp, _vp]),
    'pxb_pop_control_comb_peers': (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_double, _vp]),
    'pxb_pop_plan': (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_double, _vp]),
    'pxb_pop_pull': (ctypes.c_int, [_vp, _vp]),
    'pxb_reserve_sms': (ctypes.c_int, [_vp, ctypes.c_int]),
    'pxb_pop_control_finish': (ctypes.c_int, [_vp, _vp]),
    'pxb_bp_steps': (ctypes.c_int, [_vp]),
    'pxb_back_propagate': (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]),
    'pxb_bp_reset': (ctypes.c_int, [_vp, _vp]),
    'pxb_bp_restore_weights': (ctypes.c_int, [_vp, ctypes.c_int]),
    'pxb_bp_zero': (ctypes.c_int, [_vp, _vp]),
    'pxb_get_phi_bp': (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp]),
    'pxb_comb_plan_host': (ctypes.c_int, [_vp, ctypes.c_int64, ctypes.c_double, _vp]),
    'pxb_stage_greens': (ctypes.c_int, [_vp, ctypes.c_int, _vp]),
    'pxb_stage_force_bias_gemm': (ctypes.c_int, [_vp, _vp]),
    'pxb_get_theta': (ctypes.c_int, [_vp, _vp, _vp]),
    'pxb_get_x': (ctypes.c_int, [_vp, _vp, _vp]),
    'pxb_get_vhs': (ctypes.c_int, [_vp, _vp, _vp]),
    'pxb_get_exx': (ctypes.c_int, [_vp, _vp, _vp]),
}

_lib = None


def declared_symbols():
    return sorted(_PROTOS.keys())


def load():
    """Load the shared library and set prototypes (idempotent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "pauxy_b200: CUDA library %s not built (run `python -m pauxy_b200.build`); "
            "there is no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.pxb_abi_version() != ABI_VERSION:
        raise ImportError("pauxy_b200: ABI version mismatch")
    _lib = lib
    return lib
