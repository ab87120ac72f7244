"""Oracle: 12-D protein-docking energy.  TEST INFRASTRUCTURE ONLY.

numpy float64 restatement of `src/problem/protein_docking.py:9-48` (energy) and `:163-181` (the
loader's pair tables).  Pinned by `tests/golden/protein_*.npz`.
"""
from __future__ import annotations

import dataclasses
import os

import numpy as np

N_ATOMS = 100
DIM = 12
LB, UB = -1.5, 1.5


@dataclasses.dataclass
class ProteinInstance:
    problem_id: str
    coor_init: np.ndarray      # [100,3]
    q_vec: np.ndarray          # [100]   raw per-atom charges (file `q`)
    e_vec: np.ndarray          # [100]   raw per-atom epsilons (file `e`)
    r_vec: np.ndarray          # [100]   raw per-atom radii (file `r`)
    basis: np.ndarray          # [12,300]
    eigval: np.ndarray         # [12]
    dim: int = DIM
    lb: float = LB
    ub: float = UB
    optimum = None             # unknown (protein_docking.py:22)

    def __post_init__(self):
        # protein_docking.py:171-181 : q_ij = q_i q_j ; e_ij = sqrt(e_i e_j) ; r_ij = (r_i + r_j)/2
        q = np.tile(self.q_vec, (1, 1))
        e = np.tile(self.e_vec, (1, 1))
        r = np.tile(self.r_vec, (len(self.r_vec), 1))
        self.q = np.matmul(q.T, q)
        self.e = np.sqrt(np.matmul(e.T, e))
        self.r = (r + r.T) / 2

    def __str__(self):
        return self.problem_id


def load_instance(data_dir, problem_id) -> ProteinInstance:
    d = os.path.join(data_dir, problem_id)
    return ProteinInstance(problem_id,
                           np.loadtxt(d + '/coor_init'), np.loadtxt(d + '/q'), np.loadtxt(d + '/e'),
                           np.loadtxt(d + '/r'), np.loadtxt(d + '/basis'), np.loadtxt(d + '/eigval'))


def energy(I: ProteinInstance, x):
    """protein_docking.py:28-48.  x [NP,12] -> [NP]."""
    x = np.asarray(x, dtype=np.float64)
    inv = 1.0 / np.sqrt(I.eigval)
    coor = np.matmul(x * inv, I.basis).reshape((-1, N_ATOMS, 3)) + I.coor_init
    sq = np.expand_dims(np.sum(coor * coor, axis=-1), axis=-1)
    gram = np.matmul(coor, np.transpose(coor, (0, 2, 1)))
    d = sq - 2 * gram + np.transpose(sq, (0, 2, 1))
    d = np.sqrt(d + 0.01)
    near = (d > 0.11) & (d < 7.0)
    far = (d > 7.0) & (d < 9.0)
    d += np.eye(N_ATOMS)
    coeff = I.q / (4. * d) + np.sqrt(I.e) * ((I.r / d) ** 12 - (I.r / d) ** 6)
    return np.mean(np.sum(10 * near * coeff + 10 * far * coeff * ((9 - d) ** 2 * (-12 + 2 * d) / 8), axis=1),
                   axis=-1)


def evaluate(I: ProteinInstance, x):
    """`Basic_Problem.eval` shape dispatch (basic_problem.py:12-34)."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        return energy(I, x.reshape(1, -1))[0]
    if x.ndim == 2:
        return energy(I, x)
    return energy(I, x.reshape(-1, x.shape[-1]))


def synthetic_instance(seed, problem_id=None) -> ProteinInstance:
    """A synthetic complex with the shape/statistics of the shipped data (for benches on boxes
    where the 33 MB reference data set is absent).  Not a reference restatement."""
    g = np.random.RandomState(seed)
    coor = g.normal(size=(N_ATOMS, 3)) * 9.0
    q = g.uniform(-0.8, 0.8, size=N_ATOMS)
    e = -g.uniform(0.02, 0.25, size=N_ATOMS)      # shipped epsilons are negative (e.g. -0.2384 ... -0.0486)
    r = g.uniform(1.0, 2.2, size=N_ATOMS)
    basis = g.normal(size=(DIM, 3 * N_ATOMS)) / np.sqrt(3 * N_ATOMS)
    eig = np.sort(g.uniform(0.02, 2.0, size=DIM))
    return ProteinInstance(problem_id or f'SYN{seed}', coor, q, e, r, basis, eig)
