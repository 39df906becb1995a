"""TEST INFRASTRUCTURE ONLY — pins the replay-sampler restatement (oracle/replay_oracle.py) against the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden_replay
tools/replay.py imports `gym.spaces.Dict` only for isinstance checks; a two-line stub module stands in for gym (absent from this image).
The reference ReplayBuffer (device='cpu', save_episodes=False, temp directory) and the oracle receive the same seeded episodes
(ragged lengths, capacity small enough to evict) and are sampled under the same numpy seeds; batches must be byte-identical.  The sha256
of every reference batch is written to tests/golden/replay_small.json for the CPU/GPU tests that run without /root/reference.
"""
import contextlib
import hashlib
import io
import json
import os
import sys
import tempfile
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.replay_oracle import ReplayOracle           # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "replay_small.json")
CASES = [dict(capacity=10_000, min_t=0), dict(capacity=400, min_t=0), dict(capacity=400, min_t=3)]
L, B, N_EPISODES = 16, 7, 14


def make_episode(rng, n, A=6, img=16, clip=12):
    first = np.zeros(n, bool); first[0] = True
    last = np.zeros(n, bool); last[-1] = True
    return {"observation": rng.integers(0, 256, (n, 3, img, img), dtype=np.uint8), "action": rng.standard_normal((n, A)).astype(np.float32),
            "reward": rng.standard_normal((n, 1)).astype(np.float32), "discount": np.ones((n, 1), np.float32), "is_first": first, "is_last": last,
            "is_terminal": np.zeros(n, bool), "clip_video": rng.standard_normal((n, clip)).astype(np.float32),
            "odd": rng.standard_normal((n, 3)).astype(np.float32)}


def batch_digest(batch):
    h = hashlib.sha256()
    for k in sorted(batch):
        v = np.ascontiguousarray(np.asarray(batch[k]))
        h.update(k.encode()); h.update(str(v.dtype).encode()); h.update(str(v.shape).encode()); h.update(v.tobytes())
    return h.hexdigest()


def script(case):
    """yields ('store', episode) / ('sample', seed) operations — shared by the generator and the tests"""
    rng = np.random.default_rng(0)
    for i in range(N_EPISODES):
        yield "store", make_episode(rng, int(rng.integers(24, 120)))
        for _ in range(2):
            yield "sample", 100 + i


def import_reference_replay():
    gym = types.ModuleType("gym"); spaces = types.ModuleType("gym.spaces"); spaces.Dict = dict; gym.spaces = spaces
    sys.modules.setdefault("gym", gym); sys.modules.setdefault("gym.spaces", spaces)
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    import tools.replay as R
    return R


def main():
    R = import_reference_replay()
    out = []
    for case in CASES:
        keys = list(make_episode(np.random.default_rng(1), 24))
        spec = {k: None for k in keys}
        with tempfile.TemporaryDirectory() as d, contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            ref = R.ReplayBuffer([spec], [], d, length=L, capacity=case["capacity"], device="cpu", save_episodes=False, min_t_sampling=case["min_t"])
        ref.batch_size = B
        it = iter(ref)
        ora = ReplayOracle(L, case["capacity"], case["min_t"])
        digests = []
        for op, arg in script(case):
            if op == "store":
                ref.store_episode(episode=arg, run_checks=False)
                ora.store_episode(arg)
                assert ref._loaded_episodes == len(ora.lens) and ref._loaded_steps == ora.loaded_steps
            else:
                np.random.seed(arg)
                rb = {k: v.numpy() for k, v in next(it).items()}
                np.random.seed(arg)
                ob = ora.sample(B)
                assert set(rb) == set(ob)
                for k in rb:
                    assert rb[k].dtype == ob[k].dtype and np.array_equal(rb[k], ob[k]), (case, k)
                digests.append(batch_digest(rb))
        out.append(dict(case, digests=digests))
        print(f"[replay {case}] oracle == reference on {len(digests)} batches")
    json.dump(dict(L=L, B=B, n_episodes=N_EPISODES, cases=out, reference_commit="4cd8394", numpy=np.__version__), open(GOLDEN, "w"), indent=1)
    print("wrote", GOLDEN)


if __name__ == "__main__":
    main()
