"""Stage the UNMODIFIED reference (GMC-DRL/MetaBox) hot path under baseline/_ref/ so that it travels to the GPU box.

    python baseline/stage_reference.py [--force]

The reference is a script tree without setup.py / pyproject.toml (`pip install --target baseline/_ref /root/reference` has nothing to
install), so "installing" it = copying the packages the path imports, byte for byte, from where they lie under /root/reference:

    src/{problem,optimizer,agent,environment}/   (incl. problem/protein_docking_data, 33 MB: the 280 protein instances of config C5)
    src/utils.py, src/config.py
    src/agent_model/test/bbob_easy/{RLEPSO,LDE}_Agent.pkl                    (shipped policies)
    src/agent_model/rollout/protein_easy/RLEPSO_Agent/checkpoint20.pkl       (C5's policy)
    src/agent_model/rollout/bbob_easy/GLEET_Agent/checkpoint20.pkl
    src/agent_model/test/bbob_easy/DE_DDQN_Agent.pkl  ->  re-pickled WITHOUT its 1e5-entry replay buffer (85 MB -> < 1 MB; the buffer is
                                                         training state, rollout_episode never reads it)

baseline/_ref/ is git-ignored (reference sources never enter this repository's history) and NOT gpurun-ignored.  Nothing under
metabox_b200/ imports it; the only consumers are `bench.py --impl reference` / the `cpu_baseline` leg (baseline/ref_runner.py) and the
staged-data loaders that fail loudly when it is absent.  MANIFEST.json records the sha256 of every staged source file next to the
sha256 of its origin, so "unmodified" is checkable."""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
REF = os.environ.get('METABOX_REFERENCE', '/root/reference')
PKGS = ['problem', 'optimizer', 'agent', 'environment']
FILES = ['utils.py', 'config.py', '__init__.py']
PICKLES = ['agent_model/test/bbob_easy/RLEPSO_Agent.pkl', 'agent_model/test/bbob_easy/LDE_Agent.pkl',
           'agent_model/rollout/protein_easy/RLEPSO_Agent/checkpoint20.pkl', 'agent_model/rollout/bbob_easy/GLEET_Agent/checkpoint20.pkl']


def _sha(p):
    h = hashlib.sha256()
    with open(p, 'rb') as f:
        for blk in iter(lambda: f.read(1 << 20), b''):
            h.update(blk)
    return h.hexdigest()


def staged():
    return os.path.exists(os.path.join(DST, 'MANIFEST.json'))


def stage(force=False):
    src = os.path.join(REF, 'src')
    if not os.path.isdir(src):
        raise RuntimeError(f'reference not found at {REF} (baseline/_ref can only be staged in the build container)')
    if staged() and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(os.path.join(DST, 'src'))
    manifest = {}
    for pkg in PKGS:
        shutil.copytree(os.path.join(src, pkg), os.path.join(DST, 'src', pkg), ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    for f in FILES:
        shutil.copy(os.path.join(src, f), os.path.join(DST, 'src', f))
    for f in PICKLES:
        os.makedirs(os.path.dirname(os.path.join(DST, 'src', f)), exist_ok=True)
        shutil.copy(os.path.join(src, f), os.path.join(DST, 'src', f))
    shutil.copy(os.path.join(REF, 'LICENSE'), os.path.join(DST, 'LICENSE'))
    for root, dirs, files in os.walk(os.path.join(DST, 'src')):
        os.chmod(root, 0o755)
        for f in files:
            p = os.path.join(root, f)
            os.chmod(p, 0o644)
            if f.endswith('.py'):
                rel = os.path.relpath(p, os.path.join(DST, 'src'))
                a, b = _sha(p), _sha(os.path.join(src, rel))
                assert a == b, rel
                manifest[rel] = a
    # DE-DDQN: the shipped pickle is the whole agent incl. replay buffer -> drop the buffer, keep the agent object otherwise as it is
    sys.path.insert(0, HERE)
    import ref_runner
    ref_runner.activate(os.path.join(DST, 'src'))
    import pickle
    with open(os.path.join(src, 'agent_model/test/bbob_easy/DE_DDQN_Agent.pkl'), 'rb') as f:
        ag = pickle.load(f)
    from agent.utils import ReplayBuffer
    ag._DE_DDQN_Agent__replay_buffer = ReplayBuffer(8)
    out = os.path.join(DST, 'src/agent_model/test/bbob_easy/DE_DDQN_Agent.pkl')
    with open(out, 'wb') as f:
        pickle.dump(ag, f, -1)
    with open(os.path.join(DST, 'MANIFEST.json'), 'w') as f:
        json.dump({'reference': REF, 'python_sources_sha256': manifest,
                   'note': 'python sources byte-identical to the reference; DE_DDQN_Agent.pkl re-pickled without its replay buffer'}, f, indent=1)
    return DST


if __name__ == '__main__':
    print(stage(force='--force' in sys.argv))
