"""ctypes binding of the C ABI declared in include/iglu_b200.h.

The product path has NO CPU fallback: importing this module on a box where the CUDA library is
missing (and cannot be built) raises, and so does every entry point without a CUDA device.
"""
import ctypes as C
import os

from . import build as _build

CELLS = 1089
GRID_STRIDE = 1104
TLIST_MAX = 1096
TLIST_HDR = 8
TLIST_ROW_MAX = TLIST_HDR + TLIST_MAX

WALKING_DISCRETE, FLYING, WALKING_DICT = 0, 1, 2
REWARD_ON_CHANGE, REWARD_EVERY_STEP = 0, 1


class Config(C.Structure):
    _fields_ = [('action_space', C.c_int32), ('select_and_place', C.c_int32), ('max_steps', C.c_int32),
                ('auto_reset', C.c_int32), ('reward_mode', C.c_int32), ('write_grid_obs', C.c_int32),
                ('tlist_cap', C.c_int32), ('reserved', C.c_int32),
                ('right_placement_scale', C.c_double), ('wrong_placement_scale', C.c_double)]


class State(C.Structure):
    _fields_ = [('n', C.c_int64), ('grid', C.c_void_p), ('init', C.c_void_p), ('target', C.c_void_p),
                ('pose', C.c_void_p), ('init_pose', C.c_void_p), ('istate', C.c_void_p),
                ('tlist', C.c_void_p), ('tmeta', C.c_void_p), ('sched', C.c_void_p), ('worklist', C.c_void_p)]


class Obs(C.Structure):
    _fields_ = [('grid', C.c_void_p), ('target_grid', C.c_void_p), ('agent_pos', C.c_void_p),
                ('inventory', C.c_void_p), ('compass', C.c_void_p), ('reward', C.c_void_p),
                ('done', C.c_void_p)]


class TaskPool(C.Structure):
    _fields_ = [('m', C.c_int64), ('target', C.c_void_p), ('init', C.c_void_p), ('tlist', C.c_void_p),
                ('tmeta', C.c_void_p), ('cursor', C.c_void_p), ('next_task', C.c_void_p),
                ('seed', C.c_uint64), ('env_offset', C.c_int64)]


class RandomTaskSpec(C.Structure):
    """iglu_random_task_spec (include/iglu_b200.h)"""
    _fields_ = [('max_blocks', C.c_int32), ('height_levels', C.c_int32), ('max_dist', C.c_int32),
                ('num_colors', C.c_int32)]


class StagedTasks(C.Structure):
    """iglu_staged_tasks (include/iglu_b200.h)"""
    _fields_ = [('m', C.c_int64), ('tlist', C.c_void_p), ('tmeta', C.c_void_p), ('next_stage', C.c_void_p),
                ('stage_of', C.c_void_p), ('progressive', C.c_int32), ('reserved', C.c_int32),
                ('state', C.c_void_p), ('out', C.c_void_p)]


class EnvSnapshot(C.Structure):
    """iglu_env_snapshot (include/iglu_b200.h)"""
    _fields_ = [('pose', C.c_double * 6), ('reward', C.c_double), ('agent_pos', C.c_float * 5),
                ('inventory_obs', C.c_float * 6), ('compass', C.c_float), ('inventory', C.c_int32 * 6),
                ('active_block', C.c_int32), ('time_int_steps', C.c_int32), ('error', C.c_int32),
                ('step_no', C.c_int32), ('grid_size', C.c_int32), ('max_int', C.c_int32),
                ('target_size', C.c_int32), ('done', C.c_int32),
                ('grid', C.c_int32 * CELLS), ('target_grid', C.c_int32 * CELLS)]


EXPORTS = ('iglu_abi_version', 'iglu_strerror', 'iglu_ctx_create', 'iglu_ctx_destroy', 'iglu_set_tasks',
           'iglu_reset', 'iglu_step_walking', 'iglu_step_flying', 'iglu_step_walking_dict',
           'iglu_step_walking_host', 'iglu_max_intersection', 'iglu_last_launch_count',
           'iglu_task_pool_prepare', 'iglu_task_pool_attach', 'iglu_step_walking_host_async', 'iglu_wait',
           'iglu_fetch_env_host', 'iglu_max_intersection_strided', 'iglu_sample_tasks',
           'iglu_staged_prepare', 'iglu_staged_step', 'iglu_debug_overlap_trace', 'iglu_render_pov')

_lib = None


def library_path():
    return _build.SO


def load():
    """Loads (building first if stale/missing) libiglu_b200.so. Raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    so = _build.build_library()
    if not os.path.exists(so):
        raise RuntimeError(f'{so} is missing: the CUDA extension is required (there is no CPU fallback)')
    L = C.CDLL(so)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    L.iglu_abi_version.restype = i32
    L.iglu_strerror.restype = C.c_char_p
    L.iglu_strerror.argtypes = [i32]
    L.iglu_ctx_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.iglu_ctx_destroy.argtypes = [vp]
    L.iglu_ctx_destroy.restype = None
    L.iglu_set_tasks.argtypes = [vp, C.POINTER(State), C.POINTER(Obs), vp, vp, vp, vp, i64, vp]
    L.iglu_reset.argtypes = [vp, C.POINTER(State), C.POINTER(Obs), vp, i32, vp]
    L.iglu_step_walking.argtypes = [vp, C.POINTER(State), vp, C.POINTER(Obs), vp]
    L.iglu_step_flying.argtypes = [vp, C.POINTER(State), vp, vp, vp, vp, C.POINTER(Obs), vp]
    L.iglu_step_walking_dict.argtypes = [vp, C.POINTER(State), vp, vp, C.POINTER(Obs), vp]
    L.iglu_step_walking_host.argtypes = [vp, C.POINTER(State), vp, vp, C.POINTER(Obs), vp, vp, vp]
    L.iglu_step_walking_host_async.argtypes = [vp, C.POINTER(State), vp, vp, C.POINTER(Obs), vp, vp, vp]
    L.iglu_wait.argtypes = [vp]
    L.iglu_fetch_env_host.argtypes = [vp, C.POINTER(State), C.POINTER(Obs), i64, vp, vp, vp]
    L.iglu_max_intersection.argtypes = [vp, vp, vp, i32, vp, vp, i64, vp]
    if hasattr(L, 'iglu_max_intersection_strided'):  # absent from older A/B builds selected through IGLU_B200_SO
        L.iglu_max_intersection_strided.argtypes = [vp, i64, vp, i64, vp, i64, i32, vp, vp, i64, vp]
    if hasattr(L, 'iglu_sample_tasks'):
        L.iglu_sample_tasks.argtypes = [vp, i64, C.POINTER(RandomTaskSpec), C.c_uint64, i64, vp]
    if hasattr(L, 'iglu_staged_step'):
        L.iglu_staged_prepare.argtypes = [vp, C.POINTER(StagedTasks), vp, vp, vp]
        L.iglu_staged_step.argtypes = [vp, C.POINTER(State), C.POINTER(StagedTasks), vp, vp]
    L.iglu_last_launch_count.argtypes = [vp]
    if hasattr(L, 'iglu_render_pov'):
        L.iglu_render_pov.argtypes = [vp, C.POINTER(State), vp, i32, i32, vp, vp]
    if hasattr(L, 'iglu_debug_overlap_trace'):
        L.iglu_debug_overlap_trace.argtypes = [vp, vp]
    L.iglu_task_pool_prepare.argtypes = [vp, C.POINTER(TaskPool), vp, vp, vp]
    L.iglu_task_pool_attach.argtypes = [vp, C.POINTER(TaskPool)]
    if L.iglu_abi_version() != 1:
        raise RuntimeError('libiglu_b200.so ABI version mismatch')
    _lib = L
    return L


def check(status):
    if status == 0:
        return
    msg = load().iglu_strerror(status).decode()
    if status == -3:
        raise ValueError(msg)
    raise RuntimeError(f'iglu_b200 error {status}: {msg}')
