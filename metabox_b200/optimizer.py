"""Optimizer layer: batched engines (B lanes in lockstep on one GPU) and single-problem drop-in twins of
the reference's backbone optimizers.

Reference interface (`src/optimizer/learnable_optimizer.py:16-30`, `basic_optimizer.py:10-15`,
docs/Customization.md:125-210):
    Learnable_Optimizer.__init__(config) / init_population(problem) -> state /
    update(action, problem) -> (state, reward, is_done); public attrs fes, cost, log_index, log_interval.
    Basic_Optimizer.run_episode(problem) -> {'cost', 'fes'}

`*_Engine` classes own device state for B lanes and call the C ABI; `*_Optimizer` classes wrap an engine
with B = 1 and the reference's exact signatures / return types so `PBO_Env`, agents and testers written
for the reference run unchanged.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from .problem import InstancePack


class Learnable_Optimizer:
    def __init__(self, config):
        self.__config = config

    def init_population(self, problem):
        raise NotImplementedError

    def update(self, action, problem):
        raise NotImplementedError


class Basic_Optimizer:
    def __init__(self, config):
        self.__config = config

    def run_episode(self, problem):
        raise NotImplementedError


def _run_config(kind, NP, dim, maxFEs, n_logpoint, log_interval, use_tape, seed, lane_base=0):
    return _lib.MbxRunConfig(kind=kind, NP=NP, dim=dim, maxFEs=int(maxFEs), n_logpoint=int(n_logpoint),
                             log_interval=int(log_interval), use_tape=int(use_tape), lane_base=int(lane_base), seed=int(seed) & (2 ** 64 - 1))


def _tape_struct(cls, tape, device, int_fields=()):
    """dict of numpy/torch arrays -> (ctypes struct, keep-alive list)."""
    import torch
    st, keep = cls(), []
    if tape:
        for name, _ in cls._fields_:
            a = tape.get(name)
            if a is None:
                continue
            dt = torch.int32 if name in int_fields else torch.float64
            t = torch.as_tensor(np.ascontiguousarray(a) if isinstance(a, np.ndarray) else a).to(device=device, dtype=dt).contiguous()
            keep.append(t)
            setattr(st, name, t.data_ptr())
    return st, keep


# ======================================================================================== RLEPSO
def rlepso_pci(NP):
    """CLPSO learning-probability table, rlepso_optimizer.py:24-25."""
    idx = np.array(list(range(NP)))
    return 0.05 + 0.45 * np.exp(10 * idx / (NP - 1)) / (np.exp(10) - 1)


class RLEPSO_Engine:
    """B lanes of the RLEPSO backbone (rlepso_optimizer.py:39-263) on one GPU."""
    NP = 100
    ACTION_DIM = 35

    def __init__(self, pack: InstancePack, lane_inst, maxFEs, n_logpoint=50, log_interval=None, seed=0, NP=None,
                 lane_base=0):
        import torch
        self.pack, self.device = pack, pack.device
        self.lane_base = int(lane_base)        # global id of lane 0: keeps Philox streams placement independent
        self.NP = int(NP or RLEPSO_Engine.NP)
        self.dim = pack.dim
        self.B = len(lane_inst)
        self.maxFEs, self.n_logpoint = int(maxFEs), int(n_logpoint)
        self.log_interval = int(log_interval or maxFEs // n_logpoint)
        self.seed = int(seed)
        dev, B, NPn, D = self.device, self.B, self.NP, self.dim
        if isinstance(lane_inst, torch.Tensor):
            self.lane_inst = lane_inst.to(device=dev, dtype=torch.int32).contiguous()
        else:
            self.lane_inst = torch.as_tensor(np.asarray(lane_inst, dtype=np.int32)).to(dev)
        f32, i32 = torch.float32, torch.int32
        z = lambda *s, dt=f32: torch.zeros(*s, dtype=dt, device=dev)
        self.x, self.v, self.pbest_x = z(B, NPn, D), z(B, NPn, D), z(B, NPn, D)
        self.c_cost, self.pbest = z(B, NPn), z(B, NPn)
        self.pni = z(B, NPn, dt=i32)
        self.gbest_x, self.gbest_val = z(B, D), z(B)
        self.gbest_idx, self.no_improve, self.fes, self.log_index, self.curve_len, self.gen = (z(B, dt=i32) for _ in range(6))
        self.curve = z(B, _lib.CURVE_LEN)
        self.done = z(B, dt=torch.uint8)
        self.feature, self.reward = z(B, 1), z(B)
        self.pci = torch.from_numpy(rlepso_pci(self.NP)).to(dev)
        self._state = _lib.MbxRlepsoState(*[t.data_ptr() for t in (
            self.x, self.v, self.pbest_x, self.c_cost, self.pbest, self.pni, self.gbest_x, self.gbest_val, self.gbest_idx,
            self.no_improve, self.fes, self.log_index, self.curve_len, self.gen, self.curve, self.done)])

    def _cfg(self, use_tape):
        return _run_config(self.pack.kind, self.NP, self.dim, self.maxFEs, self.n_logpoint, self.log_interval, use_tape,
                           self.seed, self.lane_base)

    def init_population(self, tape=None):
        """-> feature [B,1] float32 (fes/maxFEs).  `tape`: dict of recorded draws (parity tests)."""
        import torch
        cfg = self._cfg(tape is not None)
        ts, keep = _tape_struct(_lib.MbxRlepsoTape, tape, self.device, ('tour',))
        with torch.cuda.device(self.device):
            rc = _lib.lib().mbx_rlepso_init(_lib.ptr(self.pack.headers), _lib.ptr(self.pack.consts), _lib.ptr(self.lane_inst),
                                            C.byref(cfg), C.byref(self._state), C.byref(ts), _lib.ptr(self.feature), self.B,
                                            _lib.stream_ptr())
        _lib.check(rc, 'mbx_rlepso_init')
        return self.feature

    def update(self, action, tape=None, dbg=None):
        """action [B,35] float32 device tensor -> (feature [B,1], reward [B], done [B] uint8), all on device.
        Never synchronises; safe to capture in a CUDA graph (tape=None)."""
        import torch
        if action.shape != (self.B, self.ACTION_DIM):
            raise AssertionError('actions size is not right!')           # rlepso_optimizer.py:113
        if action.dtype != torch.float32 or not action.is_contiguous() or action.device != self.device:
            action = action.to(device=self.device, dtype=torch.float32).contiguous()
        cfg = self._cfg(tape is not None)
        ts, keep = _tape_struct(_lib.MbxRlepsoTape, tape, self.device, ('tour',))
        with torch.cuda.device(self.device):
            rc = _lib.lib().mbx_rlepso_step(_lib.ptr(self.pack.headers), _lib.ptr(self.pack.consts), _lib.ptr(self.lane_inst),
                                            C.byref(cfg), C.byref(self._state), C.byref(ts), _lib.ptr(self.pci),
                                            _lib.ptr(action), _lib.ptr(self.feature), _lib.ptr(self.reward), _lib.ptr(dbg),
                                            self.B, _lib.stream_ptr())
        _lib.check(rc, 'mbx_rlepso_step')
        return self.feature, self.reward, self.done

    # ---- results in the reference's format -----------------------------------------------------
    def cost_curves(self, pad=True):
        """list (len B) of python lists of gbest samples; padded to 51 like tester.py:204-205."""
        cur = self.curve.cpu().numpy().astype(np.float64)
        ln = self.curve_len.cpu().numpy()
        out = []
        for b in range(self.B):
            c = list(cur[b, :ln[b]])
            if pad:
                while len(c) < _lib.CURVE_LEN:
                    c.append(c[-1])
            out.append(c)
        return out

    def load_state(self, b, **arrays):
        """Teacher forcing for parity tests: overwrite lane b's state from numpy arrays."""
        import torch
        for k, a in arrays.items():
            t = getattr(self, k)
            t[b] = torch.as_tensor(np.asarray(a)).to(device=self.device, dtype=t.dtype)

    def read_state(self, b):
        keys = ('x', 'v', 'pbest_x', 'c_cost', 'pbest', 'pni', 'gbest_x', 'gbest_val', 'gbest_idx', 'fes', 'log_index',
                'curve_len', 'curve', 'done', 'gen')
        return {k: getattr(self, k)[b].cpu().numpy() for k in keys}


class RLEPSO_Optimizer(Learnable_Optimizer):
    """Drop-in for `RLEPSO_Optimizer` (rlepso_optimizer.py:5-263): one problem, numpy in / numpy out."""

    def __init__(self, config):
        super().__init__(config)
        config.w_decay = True
        config.NP = 100
        self.__config = config
        self.fes = None
        self.cost = None
        self.log_index = None
        self.log_interval = config.log_interval
        self.name = 'EPSO'
        self._engine = None

    def init_population(self, problem):
        pack = problem._pack if getattr(problem, '_pack', None) is not None else InstancePack([problem])
        cfg = self.__config
        seed = int(np.random.randint(0, 2 ** 31 - 1)) * 2654435761 + int(np.random.randint(0, 2 ** 31 - 1))
        self._engine = RLEPSO_Engine(pack, [problem._index], cfg.maxFEs, cfg.n_logpoint, cfg.log_interval, seed, cfg.NP)
        feat = self._engine.init_population()
        self._sync()
        return feat[0].cpu().numpy().astype(np.float64)

    def _sync(self):
        e = self._engine
        self.fes = int(e.fes[0].item())
        self.log_index = int(e.log_index[0].item())
        self.cost = e.cost_curves(pad=False)[0]

    def update(self, action, problem):
        import torch
        a = torch.as_tensor(np.asarray(action, dtype=np.float32).reshape(1, -1))
        if a.shape[-1] != 35:
            raise AssertionError('actions size is not right!')
        feat, reward, done = self._engine.update(a.to(self._engine.device))
        self._sync()
        return feat[0].cpu().numpy().astype(np.float64), int(reward[0].item()), bool(done[0].item())


# ======================================================================================== Random search
class Random_search(Basic_Optimizer):
    """Drop-in for `Random_search` (random_search.py:5-59); `run_lanes` runs many episodes at once."""

    def __init__(self, config):
        super().__init__(config)
        self.__config = config
        self.__NP = int(getattr(config, 'rs_NP', 100))            # the reference hard-codes 100 (random_search.py:13)
        self.log_interval = config.log_interval
        self.cost, self.log_index = None, None

    def run_lanes(self, pack, lane_inst, seed=0, tape_u=None, lane_base=0):
        """`lane_base`: Philox lane id of lane 0 (the global lane index when a sweep is sharded over ranks)."""
        import torch
        cfg = self.__config
        B = len(lane_inst)
        dev = pack.device
        curve = torch.zeros(B, _lib.CURVE_LEN, dtype=torch.float32, device=dev)
        clen = torch.zeros(B, dtype=torch.int32, device=dev)
        fes = torch.zeros(B, dtype=torch.int32, device=dev)
        li = torch.as_tensor(np.asarray(lane_inst, dtype=np.int32)).to(dev)
        rc_cfg = _run_config(pack.kind, self.__NP, pack.dim, cfg.maxFEs, cfg.n_logpoint, cfg.log_interval, tape_u is not None, seed, int(lane_base))
        n_it, tp = 0, None
        if tape_u is not None:
            tp = torch.as_tensor(tape_u, dtype=torch.float64).to(dev).contiguous()
            n_it = tp.shape[1]
        with torch.cuda.device(dev):
            rc = _lib.lib().mbx_random_search(_lib.ptr(pack.headers), _lib.ptr(pack.consts), _lib.ptr(li), C.byref(rc_cfg),
                                              _lib.ptr(tp), n_it, _lib.ptr(curve), _lib.ptr(clen), _lib.ptr(fes), B,
                                              _lib.stream_ptr())
        _lib.check(rc, 'mbx_random_search')
        cur, ln = curve.cpu().numpy().astype(np.float64), clen.cpu().numpy()
        return [{'cost': list(cur[b, :ln[b]]), 'fes': int(f)} for b, f in enumerate(fes.cpu().numpy())]

    def run_episode(self, problem):
        problem.reset()
        pack = problem._pack if getattr(problem, '_pack', None) is not None else InstancePack([problem])
        seed = int(np.random.randint(0, 2 ** 31 - 1)) * 2654435761 + int(np.random.randint(0, 2 ** 31 - 1))
        res = self.run_lanes(pack, [problem._index], seed)[0]
        self.cost = res['cost']
        return res


# ======================================================================================== LDE
class LDE_Engine:
    """B lanes of the LDE backbone (lde_optimizer.py:132-198) on one GPU.  The reference fixes NP = 50
    (lde_optimizer.py:10, lde_agent.py:37); any NP <= 128 is accepted here (BASELINE config C4 asks for 100)."""
    NP = 50
    BINS = 5

    def __init__(self, pack: InstancePack, lane_inst, maxFEs, n_logpoint=50, log_interval=None, seed=0, NP=None,
                 lane_base=0):
        import torch
        self.pack, self.device = pack, pack.device
        self.lane_base = int(lane_base)
        self.NP = int(NP or LDE_Engine.NP)
        self.dim = pack.dim
        self.B = len(lane_inst)
        self.maxFEs, self.n_logpoint = int(maxFEs), int(n_logpoint)
        self.log_interval = int(log_interval or maxFEs // n_logpoint)
        self.seed = int(seed)
        dev, B, NPn, D = self.device, self.B, self.NP, self.dim
        if isinstance(lane_inst, torch.Tensor):
            self.lane_inst = lane_inst.to(device=dev, dtype=torch.int32).contiguous()
        else:
            self.lane_inst = torch.as_tensor(np.asarray(lane_inst, dtype=np.int32)).to(dev)
        f32, i32 = torch.float32, torch.int32
        z = lambda *s, dt=f32: torch.zeros(*s, dtype=dt, device=dev)
        self.pop, self.fit = z(B, NPn, D), z(B, NPn)
        self.hist_sum = z(B, self.BINS, dt=torch.float64)
        self.hist_cnt = z(B, dt=i32)
        self.gbest = z(B)
        self.fes, self.log_index, self.curve_len, self.gen = (z(B, dt=i32) for _ in range(4))
        self.curve = z(B, _lib.CURVE_LEN)
        self.done = z(B, dt=torch.uint8)
        self.feature, self.reward = z(B, NPn + 2 * self.BINS), z(B)
        self._state = _lib.MbxLdeState(*[t.data_ptr() for t in (
            self.pop, self.fit, self.hist_sum, self.hist_cnt, self.gbest, self.fes, self.log_index, self.curve_len,
            self.gen, self.curve, self.done)])

    def _cfg(self, use_tape):
        return _run_config(self.pack.kind, self.NP, self.dim, self.maxFEs, self.n_logpoint, self.log_interval, use_tape,
                           self.seed, self.lane_base)

    _INT_FIELDS = ('pbest_idx', 'r1', 'r2', 'jrand')

    def init_population(self, tape=None):
        """-> feature [B, NP+10] float32 (sorted normalised fitness | histogram | mean past histogram)."""
        import torch
        cfg = self._cfg(tape is not None)
        ts, keep = _tape_struct(_lib.MbxLdeTape, tape, self.device, self._INT_FIELDS)
        with torch.cuda.device(self.device):
            rc = _lib.lib().mbx_lde_init(_lib.ptr(self.pack.headers), _lib.ptr(self.pack.consts), _lib.ptr(self.lane_inst),
                                         C.byref(cfg), C.byref(self._state), C.byref(ts), _lib.ptr(self.feature), self.B,
                                         _lib.stream_ptr())
        _lib.check(rc, 'mbx_lde_init')
        return self.feature

    def update(self, action, tape=None, dbg=None):
        """action [B, 2*NP] float32 device tensor -> (feature [B,NP+10], reward [B], done [B] uint8) on device."""
        import torch
        if tuple(action.shape) != (self.B, 2 * self.NP):
            raise ValueError(f'LDE action must be [B, 2*NP] = {(self.B, 2 * self.NP)}, got {tuple(action.shape)}')
        if action.dtype != torch.float32 or not action.is_contiguous() or action.device != self.device:
            action = action.to(device=self.device, dtype=torch.float32).contiguous()
        cfg = self._cfg(tape is not None)
        ts, keep = _tape_struct(_lib.MbxLdeTape, tape, self.device, self._INT_FIELDS)
        with torch.cuda.device(self.device):
            rc = _lib.lib().mbx_lde_step(_lib.ptr(self.pack.headers), _lib.ptr(self.pack.consts), _lib.ptr(self.lane_inst),
                                         C.byref(cfg), C.byref(self._state), C.byref(ts), _lib.ptr(action),
                                         _lib.ptr(self.feature), _lib.ptr(self.reward), _lib.ptr(dbg), self.B,
                                         _lib.stream_ptr())
        _lib.check(rc, 'mbx_lde_step')
        return self.feature, self.reward, self.done

    cost_curves = RLEPSO_Engine.cost_curves
    load_state = RLEPSO_Engine.load_state

    def read_state(self, b):
        keys = ('pop', 'fit', 'hist_sum', 'hist_cnt', 'gbest', 'fes', 'log_index', 'curve_len', 'curve', 'done', 'gen')
        return {k: getattr(self, k)[b].cpu().numpy() for k in keys}


class LDE_Optimizer(Learnable_Optimizer):
    """Drop-in for `LDE_Optimizer` (lde_optimizer.py:6-198): one problem, numpy in / numpy out, BATCH_SIZE = 1
    leading axis on state / reward like the reference."""

    def __init__(self, config):
        super().__init__(config)
        config.NP = 50
        config.BINS = 5
        config.P_INI = 1
        config.P_NUM_MIN = 2
        config.P_MIN = config.P_NUM_MIN / config.NP
        self.__config = config
        self.fes = None
        self.cost = None
        self.log_index = None
        self.log_interval = config.log_interval
        self.gbest_cost = None
        self._engine = None

    def get_best(self):
        return self.gbest_cost

    def _sync(self):
        e = self._engine
        self.fes = int(e.fes[0].item())
        self.log_index = int(e.log_index[0].item())
        self.cost = e.cost_curves(pad=False)[0]
        self.gbest_cost = float(e.gbest[0].item())

    def init_population(self, problem):
        pack = problem._pack if getattr(problem, '_pack', None) is not None else InstancePack([problem])
        cfg = self.__config
        seed = int(np.random.randint(0, 2 ** 31 - 1)) * 2654435761 + int(np.random.randint(0, 2 ** 31 - 1))
        self._engine = LDE_Engine(pack, [problem._index], cfg.maxFEs, cfg.n_logpoint, cfg.log_interval, seed, cfg.NP)
        feat = self._engine.init_population()
        self._sync()
        return feat.cpu().numpy().astype(np.float64)

    def update(self, action, problem):
        import torch
        a = torch.as_tensor(np.asarray(action, dtype=np.float32).reshape(1, -1))
        feat, reward, done = self._engine.update(a.to(self._engine.device))
        self._sync()
        return feat.cpu().numpy().astype(np.float64), reward.cpu().numpy().astype(np.float64), bool(done[0].item())


# ======================================================================================== DE-DDQN
class DE_DDQN_Engine:
    """B lanes of the DE-DDQN backbone (de_ddqn_optimizer.py:46-220) on one GPU: one individual per step."""
    NP = 100
    N_FEAT = 99
    GEN_MAX, W, WS = 10, 50, 8

    def __init__(self, pack: InstancePack, lane_inst, maxFEs, n_logpoint=50, log_interval=None, seed=0, NP=None,
                 lane_base=0):
        import torch
        self.pack, self.device = pack, pack.device
        self.lane_base = int(lane_base)
        self.NP = int(NP or DE_DDQN_Engine.NP)
        self.dim = pack.dim
        self.B = len(lane_inst)
        self.maxFEs, self.n_logpoint = int(maxFEs), int(n_logpoint)
        self.log_interval = int(log_interval or maxFEs // n_logpoint)
        self.seed = int(seed)
        dev, B, NPn, D = self.device, self.B, self.NP, self.dim
        if isinstance(lane_inst, torch.Tensor):
            self.lane_inst = lane_inst.to(device=dev, dtype=torch.int32).contiguous()
        else:
            self.lane_inst = torch.as_tensor(np.asarray(lane_inst, dtype=np.int32)).to(dev)
        f32, f64, i32 = torch.float32, torch.float64, torch.int32
        z = lambda *s, dt=f32: torch.zeros(*s, dtype=dt, device=dev)
        self.X, self.cost = z(B, NPn, D), z(B, NPn)
        self.X_gbest, self.X_prebest, self.c3 = z(B, D), z(B, D), z(B, 3)
        self.ints, self.r = z(B, 8, dt=i32), z(B, 5, dt=i32)
        self.n_tot, self.n_succ = z(B, 4, self.GEN_MAX, dt=i32), z(B, 4, 4, self.GEN_MAX, dt=i32)
        self.om_sum, self.om_max = z(B, 4, 4, self.GEN_MAX, dt=f64), z(B, 4, 4, self.GEN_MAX, dt=f64)
        self.win = z(B, self.W, self.WS, dt=f64)
        self.fes, self.log_index, self.curve_len = (z(B, dt=i32) for _ in range(3))
        self.curve = z(B, _lib.CURVE_LEN)
        self.done = z(B, dt=torch.uint8)
        self.feature, self.reward = z(B, self.N_FEAT), z(B)
        self._state = _lib.MbxDdqnState(*[t.data_ptr() for t in (
            self.X, self.cost, self.X_gbest, self.X_prebest, self.c3, self.ints, self.r, self.n_tot, self.n_succ,
            self.om_sum, self.om_max, self.win, self.fes, self.log_index, self.curve_len, self.curve, self.done)])

    def _cfg(self, use_tape):
        return _run_config(self.pack.kind, self.NP, self.dim, self.maxFEs, self.n_logpoint, self.log_interval, use_tape,
                           self.seed, self.lane_base)

    def init_population(self, tape=None):
        """-> feature [B,99] float32."""
        import torch
        cfg = self._cfg(tape is not None)
        ts, keep = _tape_struct(_lib.MbxDdqnTape, tape, self.device, ('r5',))
        with torch.cuda.device(self.device):
            rc = _lib.lib().mbx_deddqn_init(_lib.ptr(self.pack.headers), _lib.ptr(self.pack.consts), _lib.ptr(self.lane_inst),
                                            C.byref(cfg), C.byref(self._state), C.byref(ts), _lib.ptr(self.feature), self.B,
                                            _lib.stream_ptr())
        _lib.check(rc, 'mbx_deddqn_init')
        return self.feature

    def update(self, action, tape=None, dbg=None):
        """action [B] int32 device tensor (0..3) -> (feature [B,99], reward [B], done [B] uint8) on device."""
        import torch
        if tuple(action.shape) != (self.B,):
            raise ValueError(f'DE-DDQN action must be [B] = {(self.B,)}, got {tuple(action.shape)}')
        if action.dtype != torch.int32 or not action.is_contiguous() or action.device != self.device:
            action = action.to(device=self.device, dtype=torch.int32).contiguous()
        cfg = self._cfg(tape is not None)
        ts, keep = _tape_struct(_lib.MbxDdqnTape, tape, self.device, ('r5',))
        with torch.cuda.device(self.device):
            rc = _lib.lib().mbx_deddqn_step(_lib.ptr(self.pack.headers), _lib.ptr(self.pack.consts), _lib.ptr(self.lane_inst),
                                            C.byref(cfg), C.byref(self._state), C.byref(ts), _lib.ptr(action),
                                            _lib.ptr(self.feature), _lib.ptr(self.reward), _lib.ptr(dbg), self.B,
                                            _lib.stream_ptr())
        _lib.check(rc, 'mbx_deddqn_step')
        return self.feature, self.reward, self.done

    def run_steps(self, packed_weights, ret, n_steps):
        """`n_steps` consecutive environment steps of every unfinished lane in ONE launch of the persistent kernel (`mbx_deddqn_run`): the
        greedy action of the Q-network (`packed_weights` = `agent.DE_DDQN_Net.pack()`, float32 on this device) and the fused step, state
        resident in shared memory.  `ret` [B] float32 accumulates the episode return in place.  -> (feature, done) device tensors."""
        import torch
        if ret.dtype != torch.float32 or tuple(ret.shape) != (self.B,) or ret.device != self.device or not ret.is_contiguous():
            raise ValueError(f'ret must be a contiguous float32 [B] = {(self.B,)} tensor on {self.device}')
        if packed_weights.dtype != torch.float32 or packed_weights.device != self.device or not packed_weights.is_contiguous():
            raise ValueError(f'packed_weights must be a contiguous float32 vector on {self.device}')
        cfg = self._cfg(False)
        with torch.cuda.device(self.device):
            rc = _lib.lib().mbx_deddqn_run(_lib.ptr(self.pack.headers), _lib.ptr(self.pack.consts), _lib.ptr(self.lane_inst), C.byref(cfg),
                                           C.byref(self._state), _lib.ptr(self.feature), _lib.ptr(packed_weights), packed_weights.numel(),
                                           _lib.ptr(ret), int(n_steps), self.B, _lib.stream_ptr())
        _lib.check(rc, 'mbx_deddqn_run')
        return self.feature, self.done

    cost_curves = RLEPSO_Engine.cost_curves
    load_state = RLEPSO_Engine.load_state

    def read_state(self, b):
        keys = ('X', 'cost', 'X_gbest', 'X_prebest', 'c3', 'ints', 'r', 'n_tot', 'n_succ', 'om_sum', 'om_max', 'win', 'fes',
                'log_index', 'curve_len', 'curve', 'done')
        return {k: getattr(self, k)[b].cpu().numpy() for k in keys}


class DE_DDQN_Optimizer(Learnable_Optimizer):
    """Drop-in for `DE_DDQN_Optimizer` (de_ddqn_optimizer.py:8-220): one problem, python int action in,
    numpy float64[99] state out."""

    def __init__(self, config):
        super().__init__(config)
        config.F = 0.5
        config.Cr = 1.0
        config.NP = 100
        config.gen_max = 10
        config.W = 50
        self.__config = config
        self.fes = None
        self.cost = None
        self.log_index = None
        self.log_interval = config.log_interval
        self._engine = None

    def _sync(self):
        e = self._engine
        self.fes = int(e.fes[0].item())
        self.log_index = int(e.log_index[0].item())
        self.cost = e.cost_curves(pad=False)[0]

    def init_population(self, problem):
        pack = problem._pack if getattr(problem, '_pack', None) is not None else InstancePack([problem])
        cfg = self.__config
        seed = int(np.random.randint(0, 2 ** 31 - 1)) * 2654435761 + int(np.random.randint(0, 2 ** 31 - 1))
        self._engine = DE_DDQN_Engine(pack, [problem._index], cfg.maxFEs, cfg.n_logpoint, cfg.log_interval, seed, cfg.NP)
        feat = self._engine.init_population()
        self._sync()
        return feat[0].cpu().numpy().astype(np.float64)

    def update(self, action, problem):
        import torch
        if int(action) not in (0, 1, 2, 3):
            raise ValueError('Action error')                              # de_ddqn_optimizer.py:153
        a = torch.tensor([int(action)], dtype=torch.int32, device=self._engine.device)
        feat, reward, done = self._engine.update(a)
        self._sync()
        return feat[0].cpu().numpy().astype(np.float64), float(reward[0].item()), bool(done[0].item())


# ======================================================================================== GLEET
class GLEET_Engine:
    """B lanes of the GLEET backbone (gleet_optimizer.py:34-314) on one GPU.  `feature` is the [B, NP, 27] state
    (observe | pbest_feature | gbest_feature per particle); the action is [B, NP] float32 in [0, 1]."""
    NP = 100
    N_FEAT = 27

    def __init__(self, pack: InstancePack, lane_inst, maxFEs, n_logpoint=50, log_interval=None, seed=0, NP=None,
                 lane_base=0):
        import torch
        self.pack, self.device = pack, pack.device
        self.lane_base = int(lane_base)
        self.NP = int(NP or GLEET_Engine.NP)
        self.dim = pack.dim
        self.B = len(lane_inst)
        self.maxFEs, self.n_logpoint = int(maxFEs), int(n_logpoint)
        self.log_interval = int(log_interval or maxFEs // n_logpoint)
        self.seed = int(seed)
        dev, B, NPn, D = self.device, self.B, self.NP, self.dim
        if isinstance(lane_inst, torch.Tensor):
            self.lane_inst = lane_inst.to(device=dev, dtype=torch.int32).contiguous()
        else:
            self.lane_inst = torch.as_tensor(np.asarray(lane_inst, dtype=np.int32)).to(dev)
        f32, i32 = torch.float32, torch.int32
        z = lambda *s, dt=f32: torch.zeros(*s, dtype=dt, device=dev)
        self.x, self.v, self.pbest_x = z(B, NPn, D), z(B, NPn, D), z(B, NPn, D)
        self.c_cost, self.pbest = z(B, NPn), z(B, NPn)
        self.pni = z(B, NPn, dt=i32)
        self.gbest_x, self.gbest_val = z(B, D), z(B)
        self.gbest_idx, self.no_improve, self.fes, self.log_index, self.curve_len, self.gen = (z(B, dt=i32) for _ in range(6))
        self.w = z(B, dt=torch.float64)
        self.max_cost = z(B)
        self.pbest_feature, self.gbest_feature = z(B, NPn, 9), z(B, 9)
        self.curve = z(B, _lib.CURVE_LEN)
        self.done = z(B, dt=torch.uint8)
        self.feature, self.reward = z(B, NPn, self.N_FEAT), z(B)
        self._state = _lib.MbxGleetState(*[t.data_ptr() for t in (
            self.x, self.v, self.pbest_x, self.c_cost, self.pbest, self.pni, self.gbest_x, self.gbest_val, self.gbest_idx,
            self.no_improve, self.fes, self.log_index, self.curve_len, self.gen, self.w, self.max_cost, self.pbest_feature,
            self.gbest_feature, self.curve, self.done)])

    def _cfg(self, use_tape):
        return _run_config(self.pack.kind, self.NP, self.dim, self.maxFEs, self.n_logpoint, self.log_interval, use_tape,
                           self.seed, self.lane_base)

    def init_population(self, tape=None):
        """-> state [B, NP, 27] float32."""
        import torch
        cfg = self._cfg(tape is not None)
        ts, keep = _tape_struct(_lib.MbxGleetTape, tape, self.device)
        with torch.cuda.device(self.device):
            rc = _lib.lib().mbx_gleet_init(_lib.ptr(self.pack.headers), _lib.ptr(self.pack.consts), _lib.ptr(self.lane_inst),
                                           C.byref(cfg), C.byref(self._state), C.byref(ts), _lib.ptr(self.feature), self.B,
                                           _lib.stream_ptr())
        _lib.check(rc, 'mbx_gleet_init')
        return self.feature

    def update(self, action, tape=None):
        """action [B, NP] float32 device tensor -> (state [B,NP,27], reward [B], done [B] uint8), all on device."""
        import torch
        if tuple(action.shape) != (self.B, self.NP):
            raise ValueError(f'GLEET action must be [B, NP] = {(self.B, self.NP)}, got {tuple(action.shape)}')
        if action.dtype != torch.float32 or not action.is_contiguous() or action.device != self.device:
            action = action.to(device=self.device, dtype=torch.float32).contiguous()
        cfg = self._cfg(tape is not None)
        ts, keep = _tape_struct(_lib.MbxGleetTape, tape, self.device)
        with torch.cuda.device(self.device):
            rc = _lib.lib().mbx_gleet_step(_lib.ptr(self.pack.headers), _lib.ptr(self.pack.consts), _lib.ptr(self.lane_inst),
                                           C.byref(cfg), C.byref(self._state), C.byref(ts), _lib.ptr(action),
                                           _lib.ptr(self.feature), _lib.ptr(self.reward), self.B, _lib.stream_ptr())
        _lib.check(rc, 'mbx_gleet_step')
        return self.feature, self.reward, self.done

    cost_curves = RLEPSO_Engine.cost_curves
    load_state = RLEPSO_Engine.load_state

    def read_state(self, b):
        keys = ('x', 'v', 'pbest_x', 'c_cost', 'pbest', 'pni', 'gbest_x', 'gbest_val', 'gbest_idx', 'no_improve', 'w',
                'max_cost', 'pbest_feature', 'gbest_feature', 'fes', 'log_index', 'curve_len', 'curve', 'done', 'gen')
        return {k: getattr(self, k)[b].cpu().numpy() for k in keys}


class GLEET_Optimizer(Learnable_Optimizer):
    """Drop-in for `GLEET_Optimizer` (gleet_optimizer.py:6-314): one problem, numpy in / numpy out."""

    def __init__(self, config):
        super().__init__(config)
        self.__config = config
        self.dim = config.dim
        self.ps = 100
        self.max_fes = config.maxFEs
        self.fes = None
        self.cost = None
        self.log_index = None
        self.log_interval = config.log_interval
        self._engine = None

    def init_population(self, problem):
        pack = problem._pack if getattr(problem, '_pack', None) is not None else InstancePack([problem])
        cfg = self.__config
        seed = int(np.random.randint(0, 2 ** 31 - 1)) * 2654435761 + int(np.random.randint(0, 2 ** 31 - 1))
        self._engine = GLEET_Engine(pack, [problem._index], cfg.maxFEs, cfg.n_logpoint, cfg.log_interval, seed, self.ps)
        state = self._engine.init_population()
        self._sync()
        return state[0].cpu().numpy().astype(np.float64)

    def _sync(self):
        e = self._engine
        self.fes = int(e.fes[0].item())
        self.log_index = int(e.log_index[0].item())
        self.cost = e.cost_curves(pad=False)[0]

    def update(self, action, problem):
        import torch
        a = torch.as_tensor(np.asarray(action, dtype=np.float32).reshape(1, -1))
        state, reward, done = self._engine.update(a.to(self._engine.device))
        self._sync()
        return state[0].cpu().numpy().astype(np.float64), float(reward[0].item()), bool(done[0].item())


# ======================================================================================== RL_PSO
class RL_PSO_Engine:
    """B lanes of the RL_PSO backbone (rl_pso_optimizer.py:32-148) on one GPU: one particle per step.  `feature` is the
    [B, 2D] state (gbest_position | current_position of the next particle); the action is [B] float32."""
    NP = 100

    def __init__(self, pack: InstancePack, lane_inst, maxFEs, n_logpoint=50, log_interval=None, seed=0, NP=None, lane_base=0):
        import torch
        self.pack, self.device = pack, pack.device
        self.lane_base = int(lane_base)
        self.NP = int(NP or RL_PSO_Engine.NP)
        self.dim = pack.dim
        self.B = len(lane_inst)
        self.maxFEs, self.n_logpoint = int(maxFEs), int(n_logpoint)
        self.log_interval = int(log_interval or maxFEs // n_logpoint)
        self.seed = int(seed)
        dev, B, NPn, D = self.device, self.B, self.NP, self.dim
        if isinstance(lane_inst, torch.Tensor):
            self.lane_inst = lane_inst.to(device=dev, dtype=torch.int32).contiguous()
        else:
            self.lane_inst = torch.as_tensor(np.asarray(lane_inst, dtype=np.int32)).to(dev)
        f32, i32 = torch.float32, torch.int32
        z = lambda *s, dt=f32: torch.zeros(*s, dtype=dt, device=dev)
        self.x, self.v, self.pbest_x = z(B, NPn, D), z(B, NPn, D), z(B, NPn, D)
        self.c_cost, self.pbest = z(B, NPn), z(B, NPn)
        self.gbest_x, self.gbest_val, self.max_cost = z(B, D), z(B), z(B)
        self.w = z(B, dt=torch.float64)
        self.gbest_idx, self.cur_index, self.fes, self.log_index, self.curve_len = (z(B, dt=i32) for _ in range(5))
        self.curve = z(B, _lib.CURVE_LEN)
        self.done = z(B, dt=torch.uint8)
        self.feature, self.reward = z(B, 2 * D), z(B)
        self._state = _lib.MbxRlPsoState(*[t.data_ptr() for t in (
            self.x, self.v, self.pbest_x, self.c_cost, self.pbest, self.gbest_x, self.gbest_val, self.max_cost, self.w,
            self.gbest_idx, self.cur_index, self.fes, self.log_index, self.curve_len, self.curve, self.done)])

    def _cfg(self, use_tape):
        return _run_config(self.pack.kind, self.NP, self.dim, self.maxFEs, self.n_logpoint, self.log_interval, use_tape,
                           self.seed, self.lane_base)

    def init_population(self, tape=None):
        import torch
        cfg = self._cfg(tape is not None)
        ts, keep = _tape_struct(_lib.MbxRlPsoTape, tape, self.device)
        with torch.cuda.device(self.device):
            rc = _lib.lib().mbx_rl_pso_init(_lib.ptr(self.pack.headers), _lib.ptr(self.pack.consts), _lib.ptr(self.lane_inst),
                                            C.byref(cfg), C.byref(self._state), C.byref(ts), _lib.ptr(self.feature), self.B,
                                            _lib.stream_ptr())
        _lib.check(rc, 'mbx_rl_pso_init')
        return self.feature

    def update(self, action, tape=None):
        """action [B] (or [B,1]) float32 device tensor -> (state [B,2D], reward [B], done [B] uint8), all on device."""
        import torch
        action = action.reshape(-1)
        if action.shape[0] != self.B:
            raise ValueError(f'RL_PSO action must have one value per lane ({self.B}), got {tuple(action.shape)}')
        if action.dtype != torch.float32 or not action.is_contiguous() or action.device != self.device:
            action = action.to(device=self.device, dtype=torch.float32).contiguous()
        cfg = self._cfg(tape is not None)
        ts, keep = _tape_struct(_lib.MbxRlPsoTape, tape, self.device)
        with torch.cuda.device(self.device):
            rc = _lib.lib().mbx_rl_pso_step(_lib.ptr(self.pack.headers), _lib.ptr(self.pack.consts), _lib.ptr(self.lane_inst),
                                            C.byref(cfg), C.byref(self._state), C.byref(ts), _lib.ptr(action),
                                            _lib.ptr(self.feature), _lib.ptr(self.reward), self.B, _lib.stream_ptr())
        _lib.check(rc, 'mbx_rl_pso_step')
        return self.feature, self.reward, self.done

    cost_curves = RLEPSO_Engine.cost_curves
    load_state = RLEPSO_Engine.load_state

    def read_state(self, b):
        keys = ('x', 'v', 'pbest_x', 'c_cost', 'pbest', 'gbest_x', 'gbest_val', 'gbest_idx', 'max_cost', 'w', 'cur_index', 'fes',
                'log_index', 'curve_len', 'curve', 'done')
        return {k: getattr(self, k)[b].cpu().numpy() for k in keys}


class RL_PSO_Optimizer(Learnable_Optimizer):
    """Drop-in for `RL_PSO_Optimizer` (rl_pso_optimizer.py:6-148): one problem, numpy in / numpy out."""

    def __init__(self, config):
        super().__init__(config)
        config.w_decay = True
        config.c = 2.05
        config.NP = 100
        self.__config = config
        self.fes = None
        self.cost = None
        self.log_index = None
        self.log_interval = config.log_interval
        self._engine = None

    def init_population(self, problem):
        pack = problem._pack if getattr(problem, '_pack', None) is not None else InstancePack([problem])
        cfg = self.__config
        seed = int(np.random.randint(0, 2 ** 31 - 1)) * 2654435761 + int(np.random.randint(0, 2 ** 31 - 1))
        self._engine = RL_PSO_Engine(pack, [problem._index], cfg.maxFEs, cfg.n_logpoint, cfg.log_interval, seed, cfg.NP)
        state = self._engine.init_population()
        self._sync()
        return state[0].cpu().numpy().astype(np.float64)

    def _sync(self):
        e = self._engine
        self.fes = int(e.fes[0].item())
        self.log_index = int(e.log_index[0].item())
        self.cost = e.cost_curves(pad=False)[0]

    def update(self, action, problem):
        import torch
        a = torch.as_tensor(np.asarray(action, dtype=np.float32).reshape(-1))
        state, reward, done = self._engine.update(a.to(self._engine.device))
        self._sync()
        return state[0].cpu().numpy().astype(np.float64), float(reward[0].item()), bool(done[0].item())


# ======================================================================================== QLPSO
class QLPSO_Engine:
    """B lanes of the QLPSO backbone (qlpso_optimizer.py:74-126) on one GPU: one particle per step.  `feature` is the [B]
    int32 Q-state (last action of the particle the next action applies to); the action is [B] int32 in 0..3."""
    NP = 30

    def __init__(self, pack: InstancePack, lane_inst, maxFEs, n_logpoint=50, log_interval=None, seed=0, NP=None, lane_base=0):
        import torch
        self.pack, self.device = pack, pack.device
        self.lane_base = int(lane_base)
        self.NP = int(NP or QLPSO_Engine.NP)
        self.dim = pack.dim
        self.B = len(lane_inst)
        self.maxFEs, self.n_logpoint = int(maxFEs), int(n_logpoint)
        self.log_interval = int(log_interval or maxFEs // n_logpoint)
        self.seed = int(seed)
        dev, B, NPn, D = self.device, self.B, self.NP, self.dim
        if isinstance(lane_inst, torch.Tensor):
            self.lane_inst = lane_inst.to(device=dev, dtype=torch.int32).contiguous()
        else:
            self.lane_inst = torch.as_tensor(np.asarray(lane_inst, dtype=np.int32)).to(dev)
        f32, i32 = torch.float32, torch.int32
        z = lambda *s, dt=f32: torch.zeros(*s, dtype=dt, device=dev)
        self.x, self.v, self.pbest_x = z(B, NPn, D), z(B, NPn, D), z(B, NPn, D)
        self.cost, self.gbest_cost = z(B, NPn), z(B)
        self.diversity = z(B, dt=torch.float64)
        self.q_state = z(B, NPn, dt=i32)
        self.pointer, self.fes, self.log_index, self.curve_len = (z(B, dt=i32) for _ in range(4))
        self.curve = z(B, _lib.CURVE_LEN)
        self.done = z(B, dt=torch.uint8)
        self.feature, self.reward = z(B, dt=i32), z(B)
        self._state = _lib.MbxQlpsoState(*[t.data_ptr() for t in (
            self.x, self.v, self.pbest_x, self.cost, self.gbest_cost, self.diversity, self.q_state, self.pointer, self.fes,
            self.log_index, self.curve_len, self.curve, self.done)])

    _cfg = RL_PSO_Engine._cfg

    @property
    def gbest_val(self):                                          # (name shared with the other engines: sweeps, Tester)
        return self.gbest_cost

    def init_population(self, tape=None):
        import torch
        cfg = self._cfg(tape is not None)
        ts, keep = _tape_struct(_lib.MbxQlpsoTape, tape, self.device, int_fields=('init_state',))
        with torch.cuda.device(self.device):
            rc = _lib.lib().mbx_qlpso_init(_lib.ptr(self.pack.headers), _lib.ptr(self.pack.consts), _lib.ptr(self.lane_inst),
                                           C.byref(cfg), C.byref(self._state), C.byref(ts), _lib.ptr(self.feature), self.B,
                                           _lib.stream_ptr())
        _lib.check(rc, 'mbx_qlpso_init')
        return self.feature

    def update(self, action, tape=None):
        """action [B] (or [B,1]) integer device tensor -> (q-state [B] int32, reward [B] float32, done [B] uint8), on device."""
        import torch
        action = action.reshape(-1)
        if action.shape[0] != self.B:
            raise ValueError(f'QLPSO action must have one value per lane ({self.B}), got {tuple(action.shape)}')
        if action.dtype != torch.int32 or not action.is_contiguous() or action.device != self.device:
            action = action.to(device=self.device, dtype=torch.int32).contiguous()
        cfg = self._cfg(tape is not None)
        ts, keep = _tape_struct(_lib.MbxQlpsoTape, tape, self.device, int_fields=('init_state',))
        with torch.cuda.device(self.device):
            rc = _lib.lib().mbx_qlpso_step(_lib.ptr(self.pack.headers), _lib.ptr(self.pack.consts), _lib.ptr(self.lane_inst),
                                           C.byref(cfg), C.byref(self._state), C.byref(ts), _lib.ptr(action),
                                           _lib.ptr(self.feature), _lib.ptr(self.reward), self.B, _lib.stream_ptr())
        _lib.check(rc, 'mbx_qlpso_step')
        return self.feature, self.reward, self.done

    cost_curves = RLEPSO_Engine.cost_curves
    load_state = RLEPSO_Engine.load_state

    def read_state(self, b):
        keys = ('x', 'v', 'pbest_x', 'cost', 'gbest_cost', 'diversity', 'q_state', 'pointer', 'fes', 'log_index', 'curve_len',
                'curve', 'done')
        return {k: getattr(self, k)[b].cpu().numpy() for k in keys}


class QLPSO_Optimizer(Learnable_Optimizer):
    """Drop-in for `QLPSO_Optimizer` (qlpso_optimizer.py:19-126): one problem, numpy in / python scalars out."""

    def __init__(self, config):
        super().__init__(config)
        config.NP = 30
        config.C = 1.49618
        config.W = 0.729844
        self.__config = config
        self.fes = None
        self.cost = None
        self.log_index = None
        self.log_interval = config.log_interval
        self._engine = None

    def init_population(self, problem):
        pack = problem._pack if getattr(problem, '_pack', None) is not None else InstancePack([problem])
        cfg = self.__config
        seed = int(np.random.randint(0, 2 ** 31 - 1)) * 2654435761 + int(np.random.randint(0, 2 ** 31 - 1))
        self._engine = QLPSO_Engine(pack, [problem._index], cfg.maxFEs, cfg.n_logpoint, cfg.log_interval, seed, cfg.NP)
        state = self._engine.init_population()
        self._sync()
        return int(state[0].item())

    _sync = RL_PSO_Optimizer._sync

    def update(self, action, problem):
        import torch
        a = torch.as_tensor(np.asarray(action, dtype=np.int32).reshape(-1))
        state, reward, done = self._engine.update(a.to(self._engine.device))
        self._sync()
        return int(state[0].item()), int(reward[0].item()), bool(done[0].item())


# ======================================================================================== DEDQN (composed step, see dedqn.py)
from .dedqn import DEDQN_Engine, DEDQN_Optimizer  # noqa: E402,F401
from .rl_hpsde import RL_HPSDE_Engine, RL_HPSDE_Optimizer  # noqa: E402,F401  (composed step, see rl_hpsde.py)
from .l2l import L2L_Engine, L2L_Optimizer  # noqa: E402,F401
from .nrlpso import NRLPSO_Engine, NRLPSO_Optimizer  # noqa: E402,F401
from .les import LES_Engine, LES_Optimizer  # noqa: E402,F401
