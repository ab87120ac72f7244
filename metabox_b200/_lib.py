"""ctypes binding of libmbx.so (include/metabox_b200.h).  There is NO fallback: if the CUDA library is
missing or a call fails, an exception is raised."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MBX_LIB selects a measurement build (libmbx_prof.so); the default is the product library
LIB_PATH = os.environ.get('MBX_LIB') or os.path.join(_HERE, 'csrc', 'libmbx.so')

MAX_DIM, MAX_NP, CURVE_LEN = 64, 128, 51
KIND_BBOB, KIND_PROTEIN = 0, 1
EVAL_ALL_STREAMING, EVAL_NONE_STREAMING = 0x100, 0x200      # promises OR-ed into mbx_eval's `kind` (include/metabox_b200.h)


class MbxError(RuntimeError):
    pass


class MbxInstance(C.Structure):
    _fields_ = [('kind', C.c_int32), ('base', C.c_int32), ('dim', C.c_int32), ('noise', C.c_int32),
                ('n_peaks', C.c_int32), ('has_optimum', C.c_int32), ('fid', C.c_int32), ('pad_', C.c_int32),
                ('lb', C.c_double), ('ub', C.c_double), ('bias', C.c_double), ('optimum', C.c_double),
                ('bh_coef', C.c_double), ('n_alpha', C.c_double), ('n_beta', C.c_double), ('n_p', C.c_double),
                ('sc', C.c_double * 8), ('off', C.c_int64)]


INSTANCE_DTYPE = np.dtype([('kind', '<i4'), ('base', '<i4'), ('dim', '<i4'), ('noise', '<i4'), ('n_peaks', '<i4'),
                           ('has_optimum', '<i4'), ('fid', '<i4'), ('pad_', '<i4'), ('lb', '<f8'), ('ub', '<f8'),
                           ('bias', '<f8'), ('optimum', '<f8'), ('bh_coef', '<f8'), ('n_alpha', '<f8'),
                           ('n_beta', '<f8'), ('n_p', '<f8'), ('sc', '<f8', (8,)), ('off', '<i8')])


class MbxRunConfig(C.Structure):
    _fields_ = [('kind', C.c_int32), ('NP', C.c_int32), ('dim', C.c_int32), ('maxFEs', C.c_int32),
                ('n_logpoint', C.c_int32), ('log_interval', C.c_int32), ('use_tape', C.c_int32), ('lane_base', C.c_int32),
                ('seed', C.c_uint64)]


_P = C.c_void_p


class MbxRlepsoState(C.Structure):
    _fields_ = [(n, _P) for n in ('x', 'v', 'pbest_x', 'c_cost', 'pbest', 'pni', 'gbest_x', 'gbest_val', 'gbest_idx',
                                  'no_improve', 'fes', 'log_index', 'curve_len', 'gen', 'curve', 'done')]


class MbxRlepsoTape(C.Structure):
    _fields_ = [(n, _P) for n in ('init_pos_u', 'init_vel_u', 'clpso_u', 'fdr_u', 'reinit_pos_u', 'reinit_vel_u',
                                  'tour', 'rand1', 'rand2', 'reinit_u', 'noise')]


class MbxLdeState(C.Structure):
    _fields_ = [(n, _P) for n in ('pop', 'fit', 'hist_sum', 'hist_cnt', 'gbest', 'fes', 'log_index', 'curve_len', 'gen',
                                  'curve', 'done')]


class MbxLdeTape(C.Structure):
    _fields_ = [(n, _P) for n in ('init_u', 'pbest_idx', 'r1', 'r2', 'cross_u', 'jrand', 'noise')]


class MbxDdqnState(C.Structure):
    _fields_ = [(n, _P) for n in ('X', 'cost', 'X_gbest', 'X_prebest', 'c3', 'ints', 'r', 'n_tot', 'n_succ', 'om_sum',
                                  'om_max', 'win', 'fes', 'log_index', 'curve_len', 'curve', 'done')]


class MbxDdqnTape(C.Structure):
    _fields_ = [(n, _P) for n in ('init_u', 'r5', 'noise')]


class MbxGleetState(C.Structure):
    _fields_ = [(n, _P) for n in ('x', 'v', 'pbest_x', 'c_cost', 'pbest', 'pni', 'gbest_x', 'gbest_val', 'gbest_idx',
                                  'no_improve', 'fes', 'log_index', 'curve_len', 'gen', 'w', 'max_cost', 'pbest_feature',
                                  'gbest_feature', 'curve', 'done')]


class MbxGleetTape(C.Structure):
    _fields_ = [(n, _P) for n in ('init_pos_u', 'init_vel_u', 'rand1', 'rand2', 'noise')]


class MbxRlPsoState(C.Structure):
    _fields_ = [(n, _P) for n in ('x', 'v', 'pbest_x', 'c_cost', 'pbest', 'gbest_x', 'gbest_val', 'max_cost', 'w', 'gbest_idx',
                                  'cur_index', 'fes', 'log_index', 'curve_len', 'curve', 'done')]


class MbxRlPsoTape(C.Structure):
    _fields_ = [(n, _P) for n in ('init_pos_u', 'init_vel_u', 'rand1', 'noise')]


class MbxQlpsoState(C.Structure):
    _fields_ = [(n, _P) for n in ('x', 'v', 'pbest_x', 'cost', 'gbest_cost', 'diversity', 'q_state', 'pointer', 'fes', 'log_index',
                                  'curve_len', 'curve', 'done')]


class MbxQlpsoTape(C.Structure):
    _fields_ = [(n, _P) for n in ('init_pos_u', 'init_state', 'rand2', 'noise')]


_lib = None

# name -> (restype, argtypes); every symbol include/metabox_b200.h declares
SIGNATURES = {
    'mbx_last_error': (C.c_char_p, []),
    'mbx_version': (C.c_int, []),
    'mbx_sizeof_instance': (C.c_int, []),
    'mbx_sizeof_runconfig': (C.c_int, []),
    'mbx_eval': (C.c_int, [C.c_int, _P, _P, _P, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64,
                           C.c_uint32, _P, _P]),
    'mbx_rlepso_init': (C.c_int, [_P, _P, _P, C.POINTER(MbxRunConfig), C.POINTER(MbxRlepsoState),
                                  C.POINTER(MbxRlepsoTape), _P, C.c_int, _P]),
    'mbx_rlepso_step': (C.c_int, [_P, _P, _P, C.POINTER(MbxRunConfig), C.POINTER(MbxRlepsoState),
                                  C.POINTER(MbxRlepsoTape), _P, _P, _P, _P, _P, C.c_int, _P]),
    'mbx_lde_init': (C.c_int, [_P, _P, _P, C.POINTER(MbxRunConfig), C.POINTER(MbxLdeState), C.POINTER(MbxLdeTape), _P,
                               C.c_int, _P]),
    'mbx_lde_step': (C.c_int, [_P, _P, _P, C.POINTER(MbxRunConfig), C.POINTER(MbxLdeState), C.POINTER(MbxLdeTape), _P, _P,
                               _P, _P, C.c_int, _P]),
    'mbx_deddqn_init': (C.c_int, [_P, _P, _P, C.POINTER(MbxRunConfig), C.POINTER(MbxDdqnState), C.POINTER(MbxDdqnTape), _P,
                                  C.c_int, _P]),
    'mbx_deddqn_step': (C.c_int, [_P, _P, _P, C.POINTER(MbxRunConfig), C.POINTER(MbxDdqnState), C.POINTER(MbxDdqnTape), _P,
                                  _P, _P, _P, C.c_int, _P]),
    'mbx_gleet_init': (C.c_int, [_P, _P, _P, C.POINTER(MbxRunConfig), C.POINTER(MbxGleetState), C.POINTER(MbxGleetTape), _P,
                                 C.c_int, _P]),
    'mbx_gleet_step': (C.c_int, [_P, _P, _P, C.POINTER(MbxRunConfig), C.POINTER(MbxGleetState), C.POINTER(MbxGleetTape), _P,
                                 _P, _P, C.c_int, _P]),
    'mbx_rl_pso_init': (C.c_int, [_P, _P, _P, C.POINTER(MbxRunConfig), C.POINTER(MbxRlPsoState), C.POINTER(MbxRlPsoTape), _P,
                                  C.c_int, _P]),
    'mbx_rl_pso_step': (C.c_int, [_P, _P, _P, C.POINTER(MbxRunConfig), C.POINTER(MbxRlPsoState), C.POINTER(MbxRlPsoTape), _P,
                                  _P, _P, C.c_int, _P]),
    'mbx_qlpso_init': (C.c_int, [_P, _P, _P, C.POINTER(MbxRunConfig), C.POINTER(MbxQlpsoState), C.POINTER(MbxQlpsoTape), _P,
                                 C.c_int, _P]),
    'mbx_qlpso_step': (C.c_int, [_P, _P, _P, C.POINTER(MbxRunConfig), C.POINTER(MbxQlpsoState), C.POINTER(MbxQlpsoTape), _P,
                                 _P, _P, C.c_int, _P]),
    'mbx_landscape_features': (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    'mbx_rlepso_policy_weight_count': (C.c_int, []),
    'mbx_rlepso_policy': (C.c_int, [_P, _P, C.c_int, _P, _P, _P, _P, C.c_int, _P]),
    'mbx_progressive_walk': (C.c_int, [_P, _P, _P, _P, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    'mbx_rl_pso_policy_weight_count': (C.c_int, [C.c_int]),
    'mbx_rl_pso_policy': (C.c_int, [_P, _P, C.c_int, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    'mbx_qlpso_policy': (C.c_int, [_P, _P, _P, _P, C.c_int, _P]),
    'mbx_deddqn_run': (C.c_int, [_P, _P, _P, C.POINTER(MbxRunConfig), C.POINTER(MbxDdqnState), _P, _P, C.c_int, _P, C.c_int, C.c_int, _P]),
    'mbx_deddqn_run_lanes_per_cta': (C.c_int, []),
    'mbx_deddqn_policy_weight_count': (C.c_int, []),
    'mbx_deddqn_policy': (C.c_int, [_P, _P, C.c_int, _P, _P, C.c_int, _P]),
    'mbx_gleet_policy_weight_count': (C.c_int, []),
    'mbx_gleet_policy': (C.c_int, [_P, _P, C.c_int, _P, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    'mbx_random_search': (C.c_int, [_P, _P, _P, C.POINTER(MbxRunConfig), _P, C.c_int, _P, _P, _P, C.c_int, _P]),
}
OPTIONAL_SIGNATURES = {'mbx_prof_read': (C.c_int, [_P, C.c_int]), 'mbx_prof_read_deddqn': (C.c_int, [_P, C.c_int]),
                       'mbx_prof_read_deddqn_run': (C.c_int, [_P, C.c_int])}     # only in libmbx_prof.so


def lib():
    """Load libmbx.so (once).  Raises MbxError when it has not been built -- no CPU fallback exists."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MbxError(f'{LIB_PATH} not found: build it with `python -m metabox_b200.build` '
                           '(the engine has no CPU fallback)')
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        for name, (res, args) in OPTIONAL_SIGNATURES.items():
            if hasattr(L, name):
                fn = getattr(L, name)
                fn.restype, fn.argtypes = res, args
        if L.mbx_sizeof_instance() != C.sizeof(MbxInstance) or L.mbx_sizeof_instance() != INSTANCE_DTYPE.itemsize:
            raise MbxError('MbxInstance ABI mismatch between libmbx.so and metabox_b200/_lib.py')
        if L.mbx_sizeof_runconfig() != C.sizeof(MbxRunConfig):
            raise MbxError('MbxRunConfig ABI mismatch between libmbx.so and metabox_b200/_lib.py')
        _lib = L
    return _lib


def check(rc, who=''):
    if rc != 0:
        raise MbxError(f'{who} failed (rc={rc}): {lib().mbx_last_error().decode()}')


def ptr(t):
    """device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
