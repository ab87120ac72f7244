"""Problem layer of the B200 engine: drop-in twins of the reference's `Basic_Problem` API
(`src/problem/basic_problem.py:4-37`) whose `eval` runs on the GPU through the C ABI (`mbx_eval`).

* `BBOB_Problem`           -- any of F1-F24 / F101-F130 (`src/problem/bbob.py:149-890`)
* `BBOB_Dataset`           -- `BBOB_Dataset.get_datasets` (`bbob.py:904-952`): same seeded draw order, so
                              instances are bit-identical to the reference's for a given numpy
* `Protein_Docking`, `Protein_Docking_Dataset` (`src/problem/protein_docking.py:9-48,121-208`)
* `InstancePack`           -- the device-side constant table (`MbxInstance[]` + float64 constants)

Host code here only *constructs constants* (numpy, float64, reference draw order).  Fitness values --
including `optimum = f(opt)` -- are always computed by the CUDA kernels; there is no CPU evaluation path.
"""
from __future__ import annotations

import os
import time

import numpy as np

from . import _lib

NOISE_KINDS = {'none': 0, 'gauss': 1, 'uniform': 2, 'cauchy': 3}
_NAMES = ['Sphere', 'Ellipsoidal', 'Rastrigin', 'Buche_Rastrigin', 'Linear_Slope', 'Attractive_Sector',
          'Step_Ellipsoidal', 'Rosenbrock_original', 'Rosenbrock_rotated', 'Ellipsoidal_high_cond', 'Discus',
          'Bent_Cigar', 'Sharp_Ridge', 'Different_Powers', 'Rastrigin_F15', 'Weierstrass', 'Schaffers',
          'Schaffers_high_cond', 'Composite_Grie_rosen', 'Schwefel', 'Gallagher_101Peaks', 'Gallagher_21Peaks',
          'Katsuura', 'Lunacek_bi_Rastrigin']
# noisy families in id order 101.. : (base shape, name stem, moderate noise?)
_NOISY = [(1, 'Sphere_moderate', True), (8, 'Rosenbrock_moderate', True), (1, 'Sphere', False),
          (8, 'Rosenbrock', False), (7, 'Step_Ellipsoidal', False), (10, 'Ellipsoidal', False),
          (14, 'Different_Powers', False), (17, 'Schaffers', False), (19, 'Composite_Grie_rosen', False),
          (21, 'Gallagher_101Peaks', False)]
_SMALL = {'bbob': (1, 5, 6, 10, 15, 20), 'bbob-noisy': (101, 105, 115, 116, 117, 119, 120, 125)}


def _householder_rotation(dim):
    """`rotate_gen` (bbob.py:11-28): consumes np.random.normal(size=dim-n+1) for n=1..dim-1."""
    acc = np.eye(dim)
    sgn = np.ones((dim,))
    for n in range(1, dim):
        m = dim - n + 1
        v = np.random.normal(size=(m,))
        sgn[n - 1] = np.sign(v[0])
        v[0] -= sgn[n - 1] * np.sqrt((v * v).sum())
        refl = np.eye(dim)
        refl[n - 1:, n - 1:] = np.eye(m) - 2. * np.outer(v, v) / (v * v).sum()
        acc = np.dot(acc, refl)
    sgn[-1] = (-1) ** (1 - (dim % 2)) * sgn.prod()
    return (sgn * acc.T).T


class Basic_Problem:
    """Same surface as the reference base class: `reset()`, `eval(x)`, attrs dim/lb/ub/optimum/T1."""
    T1 = 0

    def reset(self):
        self.T1 = 0

    def eval(self, x):
        """1-D x -> scalar, 2-D -> [NP], N-D -> flattened rows (basic_problem.py:12-34).  float64 in,
        float64 out; the time spent is accumulated in `T1` (ms) like the reference."""
        start = time.perf_counter()
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 1:
            y = self.func(x.reshape(1, -1))[0]
        elif x.ndim == 2:
            y = self.func(x)
        else:
            y = self.func(x.reshape(-1, x.shape[-1]))
        self.T1 += (time.perf_counter() - start) * 1000
        return y

    def func(self, x):
        raise NotImplementedError


class BBOB_Problem(Basic_Problem):
    """One BBOB / noisy-BBOB instance.  Constructor arguments and side effects follow the reference
    classes: (dim, shift, rotate, bias, lb, ub); extra constructor randomness is drawn from the global
    np.random in the reference's order."""

    def __init__(self, fid, dim, shift, rotate, bias, lb, ub):
        self.fid, self.dim, self.bias, self.lb, self.ub = int(fid), int(dim), bias, lb, ub
        self.FES = 0
        self.M2 = None
        self.peaks = None                       # (y, C, w) for Gallagher
        self.noise, self.n_alpha, self.n_beta, self.n_p = 0, 0., 0., 0.
        self.bh_coef = 0.
        self.sc = np.zeros(8)
        if fid >= 101:
            fam, var = divmod(fid - 101, 3)
            self.base, stem, moderate = _NOISY[fam]
            self.noise = var + 1
            self.name = stem + '_' + ('gauss', 'uniform', 'cauchy')[var]
            lo = 0.01 if moderate else 1.
            if var == 0:
                self.n_beta = lo
            elif var == 1:
                self.n_alpha = self.n_beta = lo
            else:
                self.n_alpha, self.n_p = lo, (0.05 if moderate else 0.2)
            self.bh_coef = 100.                 # NoisyProblem.boundaryHandling, bbob.py:104-105
        else:
            self.base, self.name = fid, _NAMES[fid - 1]
            self.bh_coef = {7: 1., 17: 10., 18: 10., 21: 1., 22: 1.}.get(fid, 0.)
        D, base = dim, self.base
        lin = np.linspace(0, 1, D)
        lam = lambda c: (c ** 0.5) ** np.linspace(0, 1, D)
        cv0 = np.zeros(D); cv1 = lin.copy(); cv2 = np.zeros(D)
        # ---- constructor transforms, per base shape (line numbers: problem/bbob.py) ----
        if base == 2:
            cv0 = np.power(10, 6 * np.arange(D) / (D - 1))                     # :226
        elif base == 3:
            cv0 = lam(10.)                                                     # :234
        elif base == 4:
            shift[::2] = np.abs(shift[::2])                                    # :251
            cv0 = lam(10.)
        elif base == 5:
            sg = np.sign(shift)                                                # :273-275
            sg[sg == 0.] = np.random.choice([-1., 1.], size=(sg == 0.).sum())
            shift = sg * ub
            cv0 = np.sign(shift) * (10 ** np.linspace(0, 1, D))                # :286
        elif base in (6, 13):
            rotate = np.matmul(np.matmul(_householder_rotation(D), np.diag(lam(10.))), rotate)   # :295-296,:530-531
        elif base == 7:
            rotate = np.matmul(np.diag(lam(10.)), rotate)                      # :315-317
            self.M2 = _householder_rotation(D)
            cv0 = 100 ** np.linspace(0, 1, D)                                  # :324
        elif base == 8:
            shift *= 0.75                                                      # :361-362
            rotate = np.eye(D)
            self.sc[0] = max(1., D ** 0.5 / 8.)                                # :367
        elif base == 9:
            scale = max(1., D ** 0.5 / 8.)                                     # :424-426
            self.M2 = scale * rotate
            shift = np.matmul(0.5 * np.ones(D), self.M2) / (scale ** 2)
        elif base == 10:
            cond = 1e6 if fid == 10 else 1e4                                   # :457,:466
            cv0 = cond ** (np.arange(D) / (D - 1))                             # :453
        elif base == 14:
            cv0 = 2 + 4 * np.arange(D) / max(1, D - 1)                         # :554
        elif base == 15:
            self.M2 = np.matmul(np.matmul(rotate, np.diag(lam(10.))), _householder_rotation(D))   # :590-591
        elif base == 16:
            self.M2 = np.matmul(np.matmul(rotate, np.diag(lam(0.01))), _householder_rotation(D))  # :610-611
            aK, bK = 0.5 ** np.arange(12), 3.0 ** np.arange(12)
            self.sc[0] = np.sum(aK * np.cos(np.pi * bK))                       # f0 :614
        elif base in (17, 18):
            self.M2 = np.matmul(np.diag(lam(10. if base == 17 else 1000.)), _householder_rotation(D))  # :635-636
        elif base == 19:
            scale = max(1., D ** 0.5 / 8.)                                     # :696-698
            self.M2 = scale * rotate
            shift = np.matmul(0.5 * np.ones(D) / (scale ** 2.), self.M2)
            self.sc[0] = 10. if fid == 19 else 1.                              # factor :709,:718
        elif base == 20:
            shift = 0.5 * 4.2096874633 * np.random.choice([-1., 1.], size=D)   # :745
            cv0 = lam(10.)                                                     # :754
            cv1 = 2 * np.abs(shift)                                            # :753
            cv2 = 2 * np.sign(shift)                                           # :755
        elif base in (21, 22):
            P = 101 if base == 21 else 21                                      # :770-791
            shrink, galpha = (1., 1e3) if P == 101 else (0.98, 1e6)
            y = shrink * (np.random.rand(P, D) * (ub - lb) + lb)
            y[0] = shift * shrink
            shift = y[0]
            sa = 1000 ** np.random.permutation(np.linspace(0, 1, P - 1))
            sa = np.insert(sa, obj=0, values=np.sqrt(galpha))
            Cm = np.vstack([np.random.permutation(sa[i] ** np.linspace(-0.5, 0.5, D)) for i in range(P)])
            w = np.insert(np.linspace(1.1, 9.1, P - 1), 0, 10.)
            self.peaks = (y, Cm, w)
        elif base == 23:
            rotate = np.matmul(np.matmul(_householder_rotation(D), np.diag(lam(100.))), rotate)  # :848-849
            self.sc[0] = 10 / np.power(D, 1.2)                                 # :858,:864
            self.sc[1] = 10 / D / D                                            # :865
        elif base == 24:
            mu0 = 2.5 / 5 * ub                                                 # :874-877
            shift = np.random.choice([-1., 1.], size=D) * mu0 / 2
            rotate = np.matmul(np.matmul(_householder_rotation(D), np.diag(lam(100))), rotate)
            s = 1. - 1. / (2. * np.sqrt(D + 20.) - 8.2)                        # :887-888
            self.sc[0], self.sc[1], self.sc[2] = mu0, s, -np.sqrt((mu0 ** 2 - 1) / s)
            cv2 = 2. * np.sign(shift)
        self.shift, self.rotate = shift, rotate
        self.opt = self.shift
        self.cv = (np.asarray(cv0, dtype=np.float64), np.asarray(cv1, dtype=np.float64), np.asarray(cv2, dtype=np.float64))
        self._optimum = None
        self._pack = None
        self._index = 0

    kind = _lib.KIND_BBOB

    def __str__(self):
        return self.name

    def get_optimal(self):
        return self.opt

    @property
    def optimum(self):
        """f(opt), noise free (bbob.py:40-42) -- computed on the GPU the first time it is needed."""
        if self._optimum is None:
            InstancePack([self])
        return self._optimum

    def const_block(self):
        D = self.dim
        M2 = self.M2 if self.M2 is not None else np.zeros((D, D))
        parts = [self.shift, self.rotate.reshape(-1), np.asarray(M2).reshape(-1), *self.cv]
        if self.peaks is not None:
            y, Cm, w = self.peaks
            Ry = np.matmul(y, self.rotate.T)                                    # R y_p (Gallagher: z_p = R x - R y_p)
            parts += [y.reshape(-1), Cm.reshape(-1), w, Ry.reshape(-1)]
            # float32 screening copies (mbx_bbob.cuh gallagher_max): per peak D/2 groups (C_k, C_k+1, Ry_k, Ry_k+1), then
            # ln(w_p); the group array starts on a 16-byte boundary (every instance block starts on one, see pack_host)
            P = y.shape[0]
            if (4 * D + 2 * D * D + 3 * P * D + P) % 2:
                parts.append(np.zeros(1))
            g = np.concatenate([Cm.reshape(P, D // 2, 2), Ry.reshape(P, D // 2, 2)], axis=2).astype(np.float32).reshape(-1)
            lnw = np.zeros(P + (P % 2), dtype=np.float32); lnw[:P] = np.log(w)
            parts += [g.view(np.float64), lnw.view(np.float64)]
        return np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1) for p in parts])

    def header(self, off):
        h = np.zeros((), dtype=_lib.INSTANCE_DTYPE)
        h['kind'], h['base'], h['dim'], h['noise'], h['fid'] = self.kind, self.base, self.dim, self.noise, self.fid
        h['n_peaks'] = 0 if self.peaks is None else self.peaks[0].shape[0]
        h['has_optimum'] = 1
        h['lb'], h['ub'], h['bias'] = self.lb, self.ub, self.bias
        h['optimum'] = 0. if self._optimum is None else self._optimum
        h['bh_coef'], h['n_alpha'], h['n_beta'], h['n_p'] = self.bh_coef, self.n_alpha, self.n_beta, self.n_p
        h['sc'] = self.sc
        h['off'] = off
        return h

    def func(self, x, seed=None, gen=0):
        """GPU evaluation of x [NP, dim] float64 -> float64 [NP] (incl. noise for noisy ids)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        self.FES += x.shape[0]
        pack = self._pack if self._pack is not None else InstancePack([self])
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1)) if self.noise else 0
        return pack.eval_host(self._index, x, seed=seed, gen=gen)


class Protein_Docking(Basic_Problem):
    """protein_docking.py:9-48.  Constructor takes the per-atom vectors (files `q`, `e`, `r`)."""
    n_atoms, dim, lb, ub = 100, 12, -1.5, 1.5
    kind = _lib.KIND_PROTEIN
    noise = 0

    def __init__(self, coor_init, q, e, r, basis, eigval, problem_id):
        self.coor_init = np.asarray(coor_init, dtype=np.float64)
        self.q_vec, self.e_vec, self.r_vec = (np.asarray(a, dtype=np.float64).reshape(-1) for a in (q, e, r))
        self.basis, self.eigval = np.asarray(basis, dtype=np.float64), np.asarray(eigval, dtype=np.float64)
        self.problem_id = problem_id
        self.optimum = None
        self._pack, self._index = None, 0

    def __str__(self):
        return self.problem_id

    def const_block(self):
        # sqrt(e_ij) with e_ij = sqrt(e_i e_j) (protein_docking.py:179, :42)  ==  |e_i|^(1/4) * |e_j|^(1/4);
        # the shipped epsilons are all negative, so e_i e_j > 0 (mixed signs would be NaN in the reference too)
        if not (np.all(self.e_vec >= 0) or np.all(self.e_vec <= 0)):
            raise ValueError('protein epsilons of mixed sign: sqrt(e_i*e_j) is undefined')
        blk = np.concatenate([self.coor_init.reshape(-1), self.basis.reshape(-1), 1.0 / np.sqrt(self.eigval),
                              self.q_vec, np.sqrt(np.sqrt(np.abs(self.e_vec))), self.r_vec])
        pairs = self.candidate_pairs()
        lst = np.zeros((len(pairs) + 3) // 4 * 4, dtype=np.uint16)
        lst[:len(pairs)] = (pairs[:, 0].astype(np.uint16) << 8) | pairs[:, 1].astype(np.uint16)
        return np.concatenate([blk, lst.view(np.float64)])

    def candidate_pairs(self):
        """Atom pairs (i < j) that can come within the 9 A cut-off of protein_docking.py:36-37 for SOME x in [lb, ub]^12:
        initial distance minus the maximal displacement of both atoms (sum_k ub / sqrt(eig_k) * |basis_k,atom|) <= 9.  For
        every in-bounds candidate the reference masks the other pairs out, so the kernel may skip them (exactly)."""
        c = self.coor_init.reshape(self.n_atoms, 3)
        B = self.basis.reshape(self.dim, self.n_atoms, 3)
        disp = np.sum(max(abs(self.lb), abs(self.ub)) / np.sqrt(self.eigval)[:, None] * np.linalg.norm(B, axis=2), axis=0)
        d0 = np.linalg.norm(c[:, None, :] - c[None, :, :], axis=2)
        i, j = np.triu_indices(self.n_atoms, 1)
        keep = d0[i, j] - disp[i] - disp[j] <= 9.0 + 1e-6
        return np.stack([i[keep], j[keep]], axis=1)

    def header(self, off):
        h = np.zeros((), dtype=_lib.INSTANCE_DTYPE)
        h['kind'], h['dim'], h['lb'], h['ub'], h['off'] = self.kind, self.dim, self.lb, self.ub, off
        h['n_peaks'] = len(self.candidate_pairs())                # protein: number of candidate pairs after the block
        return h

    def func(self, x, seed=0, gen=0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        pack = self._pack if self._pack is not None else InstancePack([self])
        return pack.eval_host(self._index, x)


class InstancePack:
    """Device table of problem constants for a list of problems of one kind and one dim.
    Computes each BBOB instance's `optimum` on the GPU (float64 kernel) at construction."""

    def __init__(self, problems, device=None):
        import torch
        if not torch.cuda.is_available():
            raise _lib.MbxError('InstancePack needs a CUDA device: fitness is only ever computed by the CUDA kernels')
        _lib.lib()
        self.problems = list(problems)
        kinds = {p.kind for p in self.problems}
        dims = {p.dim for p in self.problems}
        if len(kinds) != 1 or len(dims) != 1:
            raise ValueError('InstancePack: all problems must share kind and dim')
        self.kind, self.dim = kinds.pop(), dims.pop()
        self.device = torch.device(device if device is not None else f'cuda:{torch.cuda.current_device()}')
        self.h_headers, self.h_consts = pack_host(self.problems)
        self._set_eval_kind()
        self.consts = torch.from_numpy(self.h_consts).to(self.device)
        self._upload_headers()
        if self.kind == _lib.KIND_BBOB and any(p._optimum is None for p in self.problems):
            opt = np.stack([p.opt for p in self.problems])[:, None, :]              # [n,1,D]
            lanes = np.arange(len(self.problems), dtype=np.int32)
            f = self.eval_lanes(lanes, opt, subtract_optimum=False, noise_free=True)
            for k, p in enumerate(self.problems):
                p._optimum = float(f[k, 0])
                self.h_headers[k]['optimum'] = p._optimum
            self._upload_headers()
        # A problem object is bound to the FIRST pack that holds it (`problem.eval`, `*_Optimizer.init_population(problem)` route through
        # `_pack` / `_index`): a later, larger pack over the same objects (Tester.test_for_random_search packs train + test) must not
        # silently re-route them to another table / index / device.  Every pack keeps its own index map.
        self.index_of = {id(p): k for k, p in enumerate(self.problems)}
        for k, p in enumerate(self.problems):
            if getattr(p, '_pack', None) is None:
                p._pack, p._index = self, k

    @classmethod
    def from_pinned(cls, problems, h_headers_u8, h_consts, device):
        """Upload an already packed table (pinned host tensors: headers as uint8 bytes, float64 constants)
        with asynchronous H2D copies on the current stream; no evaluation, no synchronisation."""
        self = cls.__new__(cls)
        self.problems = list(problems)
        self.kind, self.dim = self.problems[0].kind, self.problems[0].dim
        self.device = device
        # host copies stay available (eval_lanes(noise_free=True) and Tester.cal_t1 patch the headers on the host)
        self.h_headers = np.frombuffer(h_headers_u8.cpu().numpy().tobytes(), dtype=_lib.INSTANCE_DTYPE).copy()
        self.h_consts = h_consts.cpu().numpy() if hasattr(h_consts, 'cpu') else np.asarray(h_consts)
        self.index_of = {id(p): k for k, p in enumerate(self.problems)}
        self._set_eval_kind()
        self.headers = h_headers_u8.to(device, non_blocking=True)
        self.consts = h_consts.to(device, non_blocking=True)
        return self

    def _set_eval_kind(self):
        """`kind` argument of mbx_eval with the pack-wide promise about its lanes (include/metabox_b200.h): when every / no instance of the
        pack is one of the no-rotation functions (base shapes 5, 8, 20), the kernel that would skip all lanes is not launched."""
        self.eval_kind = self.kind
        if self.kind == _lib.KIND_BBOB and self.h_headers is not None:
            light = np.isin(np.asarray(self.h_headers['base']), (5, 8, 20))
            if light.all():
                self.eval_kind = self.kind | _lib.EVAL_ALL_STREAMING
            elif not light.any():
                self.eval_kind = self.kind | _lib.EVAL_NONE_STREAMING

    def _upload_headers(self):
        import torch
        raw = np.frombuffer(self.h_headers.tobytes(), dtype=np.uint8).copy()
        self.headers = torch.from_numpy(raw).to(self.device)

    def __len__(self):
        return len(self.problems)

    def eval_lanes(self, lane_inst, x, subtract_optimum=False, seed=0, gen=0, noise_free=False, noise_tape=None):
        """x: numpy/torch [B, NP, D] (float32 or float64) -> same-width costs [B, NP] (numpy)."""
        import torch
        xt = torch.as_tensor(x)
        if xt.dtype not in (torch.float32, torch.float64):
            xt = xt.to(torch.float64)
        xt = xt.to(self.device).contiguous()
        B, NP, D = xt.shape
        if D != self.dim:
            raise ValueError(f'x has dim {D}, pack has dim {self.dim}')
        li = torch.as_tensor(np.asarray(lane_inst, dtype=np.int32)).to(self.device)
        out = torch.empty((B, NP), dtype=xt.dtype, device=self.device)
        hdr = self.headers
        if noise_free and self.kind == _lib.KIND_BBOB:
            hh = self.h_headers.copy()
            hh['noise'] = 0
            hdr = torch.from_numpy(np.frombuffer(hh.tobytes(), dtype=np.uint8).copy()).to(self.device)
        nt = None if noise_tape is None else torch.as_tensor(noise_tape, dtype=torch.float64).to(self.device).contiguous()
        with torch.cuda.device(self.device):
            rc = _lib.lib().mbx_eval(self.eval_kind, _lib.ptr(hdr), _lib.ptr(self.consts), _lib.ptr(li), _lib.ptr(xt),
                                     int(xt.dtype == torch.float64), _lib.ptr(out), B, NP, D, int(subtract_optimum),
                                     seed, gen, _lib.ptr(nt), _lib.stream_ptr())
            _lib.check(rc, 'mbx_eval')
            return out.cpu().numpy()

    def eval_device(self, lane_inst, x, subtract_optimum=False, seed=0, gen=0):
        """Device-resident twin of `eval_lanes`: lane_inst int32 [B] and x float32/float64 [B, NP, D] are tensors on the pack's device,
        the result [B, NP] (same dtype as x) stays there; launched on the current stream, no synchronisation."""
        import torch
        if x.device != self.device or lane_inst.device != self.device or lane_inst.dtype != torch.int32:
            raise ValueError('eval_device: lane_inst (int32) and x must live on the pack\'s device')
        if x.dtype not in (torch.float32, torch.float64) or x.dim() != 3 or x.shape[2] != self.dim or lane_inst.shape != x.shape[:1]:
            raise ValueError(f'eval_device: x must be float32/float64 [B, NP, {self.dim}] with lane_inst [B]: got {x.dtype} {tuple(x.shape)}')
        x = x.contiguous()
        B, NP, D = x.shape
        out = torch.empty((B, NP), dtype=x.dtype, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.lib().mbx_eval(self.eval_kind, _lib.ptr(self.headers), _lib.ptr(self.consts), _lib.ptr(lane_inst), _lib.ptr(x),
                                     int(x.dtype == torch.float64), _lib.ptr(out), B, NP, D, int(subtract_optimum), seed, gen, None,
                                     _lib.stream_ptr())
        _lib.check(rc, 'mbx_eval')
        return out

    def eval_host(self, index, x, seed=0, gen=0):
        return self.eval_lanes([index], x[None], seed=seed, gen=gen)[0]


def pack_host(problems):
    """Host-side packing (no CUDA needed): returns (headers structured array, float64 constants)."""
    blocks, headers, off = [], [], 0
    for p in problems:
        blk = p.const_block()
        if blk.size % 2:                                        # keep every block 16-byte aligned
            blk = np.concatenate([blk, np.zeros(1)])
        headers.append(p.header(off))
        blocks.append(blk)
        off += blk.size
    return np.array(headers, dtype=_lib.INSTANCE_DTYPE), np.concatenate(blocks)


class _Dataset:
    """List-like container with the reference's Dataset surface (bbob.py:893-971)."""

    def __init__(self, data, batch_size=1):
        self.data, self.batch_size = list(data), batch_size
        self.N = len(self.data)
        self.ptr = list(range(0, self.N, batch_size))
        self.index = np.arange(self.N)

    def __getitem__(self, item):
        if self.batch_size < 2:
            return self.data[self.index[item]]
        p = self.ptr[item]
        return [self.data[i] for i in self.index[p: min(p + self.batch_size, self.N)]]

    def __len__(self):
        return self.N

    def __add__(self, other):
        return type(self)(self.data + other.data, self.batch_size)

    def shuffle(self):
        self.index = np.random.permutation(self.N)


class BBOB_Dataset(_Dataset):
    @staticmethod
    def get_datasets(suit, dim, upperbound, shifted=True, rotated=True, biased=True, train_batch_size=1,
                     test_batch_size=1, difficulty='easy', instance_seed=3849):
        """bbob.py:904-952 -- same arguments, same seeded draw order, same easy/difficult split."""
        if suit == 'bbob':
            ids = range(1, 25)
        elif suit == 'bbob-noisy':
            ids = range(101, 131)
        else:
            raise ValueError(f'{suit} function suit is invalid or is not supported yet.')
        if difficulty != 'easy' and difficulty != 'difficult':
            raise ValueError(f'{difficulty} difficulty is invalid.')
        if instance_seed > 0:
            np.random.seed(instance_seed)
        assert upperbound >= 5., f'Argument upperbound must be at least 5, but got {upperbound}.'
        ub, lb = upperbound, -upperbound
        train, test = [], []
        for fid in ids:
            shift = 0.8 * (np.random.random(dim) * (ub - lb) + lb) if shifted else np.zeros(dim)
            H = _householder_rotation(dim) if rotated else np.eye(dim)
            bias = np.random.randint(1, 26) * 100 if biased else 0
            inst = BBOB_Problem(fid, dim, shift, H, bias, lb, ub)
            small = fid in _SMALL[suit]
            ((train if not small else test) if difficulty == 'easy' else (train if small else test)).append(inst)
        return BBOB_Dataset(train, train_batch_size), BBOB_Dataset(test, test_batch_size)


class Protein_Docking_Dataset(_Dataset):
    proteins_set = {'rigid': ['1AVX', '1BJ1', '1BVN', '1CGI', '1DFJ', '1EAW', '1EWY', '1EZU', '1IQD', '1JPS',
                              '1KXQ', '1MAH', '1N8O', '1PPE', '1R0R', '2B42', '2I25', '2JEL', '7CEI', '1AY7'],
                    'medium': ['1GRN', '1IJK', '1M10', '1XQS', '2HRK'],
                    'difficult': ['1ATN', '1IBR', '2C0L']}
    n_start_points = 10

    @staticmethod
    def split(difficulty='easy', dataset_seed=1035):
        """Train/test complex names (protein_docking.py:137-154)."""
        ratio = {'easy': 0.75, 'difficult': 0.25}.get(difficulty)
        if ratio is None:
            raise ValueError
        if dataset_seed > 0:
            np.random.seed(dataset_seed)
        train, test = [], []
        for key in Protein_Docking_Dataset.proteins_set:
            perm = np.random.permutation(Protein_Docking_Dataset.proteins_set[key])
            n = max(1, min(int(len(perm) * ratio), len(perm) - 1))
            train.extend(perm[:n]); test.extend(perm[n:])
        return [str(t) for t in train], [str(t) for t in test]

    @staticmethod
    def get_datasets(version='protein', train_batch_size=1, test_batch_size=1, difficulty='easy', dataset_seed=1035,
                     data_folder=None):
        """protein_docking.py:130-195.  `data_folder` must point at a `protein_docking_data` directory in
        the reference's text format (280 sub-directories); it is not shipped with this repo."""
        if version != 'protein':
            raise ValueError(f'{version} version is invalid or is not supported yet.')
        data_folder = data_folder or os.environ.get('MBX_PROTEIN_DATA')
        if not data_folder or not os.path.isdir(data_folder):
            raise FileNotFoundError('protein_docking_data folder not found; pass data_folder= or set MBX_PROTEIN_DATA')
        train, test = Protein_Docking_Dataset.split(difficulty, dataset_seed)
        data = []
        for name in train + test:
            for j in range(Protein_Docking_Dataset.n_start_points):
                pid = f'{name}_{j + 1}'
                d = os.path.join(data_folder, pid)
                data.append(Protein_Docking(np.loadtxt(d + '/coor_init'), np.loadtxt(d + '/q'), np.loadtxt(d + '/e'),
                                            np.loadtxt(d + '/r'), np.loadtxt(d + '/basis'), np.loadtxt(d + '/eigval'), pid))
        n_tr = len(train) * Protein_Docking_Dataset.n_start_points
        return Protein_Docking_Dataset(data[:n_tr], train_batch_size), Protein_Docking_Dataset(data[n_tr:], test_batch_size)


def synthetic_protein(seed, problem_id=None):
    """Synthetic complex with the shape of the shipped data (for benchmarks without the data set)."""
    g = np.random.RandomState(seed)
    return Protein_Docking(g.normal(size=(100, 3)) * 9.0, g.uniform(-0.8, 0.8, size=100), -g.uniform(0.02, 0.25, size=100),
                           g.uniform(1.0, 2.2, size=100), g.normal(size=(12, 300)) / np.sqrt(300),
                           np.sort(g.uniform(0.02, 2.0, size=12)), problem_id or f'SYN{seed}')
