"""CPU: host-side logic of the product (instance construction, packing, C-ABI surface).  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

from metabox_b200 import _lib, problem as mp

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'metabox_b200.h')).read()
    declared = set(re.findall(r'\b(mbx_[a-z0-9_]+)\s*\(', hdr))
    assert declared, 'no declarations parsed'
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f'{name} declared in include/metabox_b200.h but not exported'
    bound = set(_lib.SIGNATURES) | set(_lib.OPTIONAL_SIGNATURES)
    assert declared <= bound, f'symbols without a ctypes signature: {declared - bound}'
    lib = _lib.lib()
    assert lib.mbx_version() >= 100
    assert lib.mbx_sizeof_instance() == _lib.INSTANCE_DTYPE.itemsize


def test_c_abi_rejects_bad_arguments_without_a_gpu():
    lib = _lib.lib()
    rc = lib.mbx_eval(0, None, None, None, None, 0, None, 1, 1, 10, 0, 0, 0, None, None)
    assert rc == -1 and b'null' in lib.mbx_last_error()
    cfg = _lib.MbxRunConfig(kind=0, NP=3, dim=10, maxFEs=100, n_logpoint=50, log_interval=2, use_tape=0, seed=1)
    rc = lib.mbx_rlepso_init(None, None, None, ctypes.byref(cfg), None, None, None, 1, None)
    assert rc == -1 and b'NP' in lib.mbx_last_error()


@pytest.mark.parametrize('dim', [10])
def test_dataset_matches_reference_constants(dim):
    G = np.load(os.path.join(GOLDEN, f'bbob_d{dim}.npz'))
    for suit in ('bbob', 'bbob-noisy'):
        tr, te = mp.BBOB_Dataset.get_datasets(suit, dim, 5, difficulty='easy')
        assert [str(p) for p in te.data] == list(G[f'split_easy_{suit}'])
        for p in tr.data + te.data:
            fid = p.fid
            assert str(G[f'name_{fid}']) == str(p)
            np.testing.assert_array_equal(G[f'shift_{fid}'], p.shift)
            np.testing.assert_array_equal(G[f'R_{fid}'], p.rotate)
            assert float(G[f'bias_{fid}']) == p.bias
            if f'M2_{fid}' in G.files:
                np.testing.assert_array_equal(G[f'M2_{fid}'], p.M2)
            if f'y_{fid}' in G.files:
                for k, a in zip(('y', 'C', 'w'), p.peaks):
                    np.testing.assert_array_equal(G[f'{k}_{fid}'], a)
        _, ted = mp.BBOB_Dataset.get_datasets(suit, dim, 5, difficulty='difficult')
        assert [str(p) for p in ted.data] == list(G[f'split_difficult_{suit}'])


def test_pack_layout():
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', 10, 5)
    probs = tr.data + te.data
    H, Cst = mp.pack_host(probs)
    D = 10
    for h, p in zip(H, probs):
        off = int(h['off'])
        np.testing.assert_array_equal(Cst[off:off + D], p.shift)
        np.testing.assert_array_equal(Cst[off + D:off + D + D * D].reshape(D, D), p.rotate)
        assert h['base'] == p.base and h['dim'] == D and h['fid'] == p.fid
        if p.peaks is not None:
            o = off + 4 * D + 2 * D * D
            np.testing.assert_array_equal(Cst[o:o + p.peaks[0].size].reshape(-1, D), p.peaks[0])
    assert Cst.dtype == np.float64


def test_protein_split_matches_reference():
    G = np.load(os.path.join(GOLDEN, 'protein.npz'))
    tr, te = mp.Protein_Docking_Dataset.split('easy')
    assert [f'{n}_{j + 1}' for n in tr for j in range(10)] == list(G['split_easy_train'])
    assert [f'{n}_{j + 1}' for n in te for j in range(10)] == list(G['split_easy_test'])


def test_dataset_container_api():
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', 10, 5, train_batch_size=4)
    assert len(tr) == 18 and len(te) == 6 and len(tr + te) == 24
    assert len(tr[0]) == 4
    with pytest.raises(ValueError):
        mp.BBOB_Dataset.get_datasets('cec', 10, 5)
    with pytest.raises(ValueError):
        mp.BBOB_Dataset.get_datasets('bbob', 10, 5, difficulty='hard')


def test_no_cpu_fallback_for_eval():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    tr, _ = mp.BBOB_Dataset.get_datasets('bbob', 10, 5)
    with pytest.raises(_lib.MbxError):
        tr.data[0].eval(np.zeros(10))
