"""CPU-side checks (no GPU compute): the C-ABI library loads and exports every symbol include/genrl_b200.h declares,
the agent mirrors reproduce the reference's parameter names and seeded initial weights, the product refuses CPU tensors
(no fallback), and the data-parallel host logic works over gloo with world_size 2."""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from genrl_b200._lib import lib, parse_header, LIB_PATH, GenrlError
from genrl_b200.config import genrl_config, make_specs, tiny_overrides, AttrDict
from tests.golden_utils import load_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_header_symbol():
    from genrl_b200 import build
    path = build.build()
    assert os.path.exists(path) and path == LIB_PATH
    dll = ctypes.CDLL(path)
    protos = parse_header()
    assert len(protos) >= 35
    for name in protos:
        assert hasattr(dll, name), f"{name} declared in include/genrl_b200.h but not exported"
    lib.load()
    assert lib.genrl_abi_version() == 1
    assert lib.genrl_pack_entry_size() == 48


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """the GEMM really is tcgen05 + TMA (SASS mnemonics from B200_PROFILING.md)"""
    import subprocess
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
    # round 2: 2-CTA tensor-core instructions (cta_group::2), 1-D bulk copies of the LayerNorm pipelines, cluster barriers of the chain kernel
    assert "UTCHMMA.2CTA" in sass and "UBLKCP" in sass and "UCGABAR" in sass


def test_streaming_scan_eligibility_is_a_pure_shape_query():
    """genrl_rssm_scan_stream_supported launches nothing: BASELINE configs[4]'s shard (16 sequences, deter 4096) is eligible, more than 16
    sequences, widths that are not multiples of 128 and widths needing more than 160 CTAs are not"""
    f = lib.load().genrl_rssm_scan_stream_supported
    assert f(16, 4096) == 1 and f(1, 1280) == 1 and f(16, 2048) == 1
    assert f(17, 4096) == 0 and f(16, 4000) == 0 and f(0, 4096) == 0 and f(16, 8192) == 0


@pytest.mark.parametrize("case", ["tiny_genrl", "tiny_dreamer"])
def test_agent_mirror_has_reference_parameters_and_seeded_init(case):
    from genrl_b200.agent.genrl import GenRLAgent
    from genrl_b200.agent.dreamer import DreamerAgent
    g, meta, _ = load_case(case)
    cfg = genrl_config(device="cpu", precision=32, batch_size=meta["B"], batch_length=meta["T"], imag_horizon=meta["H"],
                       agent=meta["kind"], **meta["overrides"])
    obs_space, act_spec = make_specs(action_dim=meta["A"], clip_dim=meta["clip"])
    torch.manual_seed(1)
    cls = GenRLAgent if meta["kind"] == "genrl" else DreamerAgent
    agent = cls(name=meta["kind"], cfg=cfg, obs_space=obs_space, act_spec=act_spec)
    sd = agent.state_dict()
    assert set(sd) == set(g["init"]), set(sd) ^ set(g["init"])
    for k, v in g["init"].items():
        assert torch.equal(sd[k], v), f"{k}: initial weights differ from the reference under the same seed"


def test_no_cpu_fallback():
    from genrl_b200 import ops
    x = torch.zeros(4, 8)
    with pytest.raises(GenrlError):
        ops.gemm(x, x)
    with pytest.raises(GenrlError):
        ops.ln_act(x, torch.ones(8), torch.zeros(8))


def test_attrdict_behaves_like_the_reference_cfg():
    c = genrl_config(device="cpu", **tiny_overrides())
    assert c.rssm.hidden == 64 and c.get("nope", 3) == 3 and getattr(c, "nope", 5) == 5
    c.update(foo=dict(bar=1))
    assert c.foo.bar == 1 and dict(**c.kl)["free"] == 1.0
    import copy, pickle
    assert pickle.loads(pickle.dumps(copy.deepcopy(c))).rssm.deter == 64


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from genrl_b200.dist import DataParallel
    dp = DataParallel()
    g = [torch.full((10,), float(rank + 1)), torch.arange(4, dtype=torch.float32) * (rank + 1)]
    dp.allreduce(g)
    ok = torch.allclose(g[0], torch.full((10,), 1.5)) and torch.allclose(g[1], torch.arange(4, dtype=torch.float32) * 1.5)
    gathered = dp.gather(torch.full((3,), float(rank)))
    ok = ok and torch.equal(gathered, torch.tensor([0., 0., 0., 1., 1., 1.]))
    data = {"a": torch.arange(8).reshape(4, 2), "b": torch.arange(4)}
    sh = dp.shard(data)
    ok = ok and torch.equal(sh["b"], torch.arange(4)[rank * 2:rank * 2 + 2]) and sh["a"].shape == (2, 2)
    # mean-of-shard-means == global mean (the property the sharded losses rely on)
    x = torch.arange(8, dtype=torch.float32)
    m = [x[rank * 4:(rank + 1) * 4].mean().reshape(1)]
    dp.allreduce(m)
    ok = ok and abs(float(m[0]) - float(x.mean())) < 1e-6
    q.put((rank, bool(ok), dp.n_allreduce))
    dist.destroy_process_group()


def test_data_parallel_host_logic_gloo_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(n == 3 for _, _, n in res)


def test_product_synthetic_batches_equal_the_test_side_generator():
    from genrl_b200.synthetic import synthetic_batch
    from oracle.genrl_oracle import synthetic_batch as oracle_batch
    a, b = synthetic_batch(3, 8, 5, clip_dim=16, seed=4, resets=0.2), oracle_batch(3, 8, 5, clip_dim=16, seed=4, resets=0.2)
    assert set(a) == set(b)
    for k in a:
        assert a[k].dtype == b[k].dtype and torch.equal(a[k], b[k]), k


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under genrl_b200/ may import it, and bench.py only in its CPU-baseline legs"""
    import re
    pkg = os.path.join(ROOT, "genrl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f"{f} imports oracle/"
    bench = open(os.path.join(ROOT, "bench.py")).read()
    native = bench[bench.index("def run_native"):bench.index("def roofline_pass")]
    assert "oracle" not in native.replace("cpu_baseline", "")


def test_bench_roofline_pass_is_single_gpu_only():
    """the roofline pass re-runs eager steps on rank 0 only; under data parallelism those steps contain NCCL collectives the other
    ranks never join (it deadlocked an 8-GPU run) — it must be gated on world == 1, like the CPU baseline"""
    src = open(os.path.join(ROOT, "bench.py")).read()
    call = src.index("roofline_pass(step, devb, nb, lib, args, agent, ms_step)")
    guard = src.rfind("\n    if ", 0, call)
    assert "world == 1" in src[guard:call], src[guard:call]
    cb = src.index('line["cpu_baseline"] = cpu_baseline(args, agent)')
    guard = src.rfind("\n    if ", 0, cb)
    assert "world == 1" in src[guard:cb]


def test_header_and_sources_declare_the_same_entry_points():
    """every extern "C" genrl_* function defined in csrc/ is declared in include/genrl_b200.h and vice versa (argument counts match)"""
    import re
    csrc = os.path.join(ROOT, "genrl_b200", "csrc")
    defined = {}
    for f in sorted(os.listdir(csrc)):
        if f.endswith(".cu"):
            text = open(os.path.join(csrc, f)).read()
            for m in re.finditer(r'extern "C"\s+(?:const char\*|int|long long)\s+(genrl_\w+)\s*\(([^)]*)\)\s*\{', text, re.S):
                args = m.group(2).strip()
                defined[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    protos = parse_header()
    assert set(defined) == set(protos), set(defined) ^ set(protos)
    for name, (_, argtypes, _) in protos.items():
        assert len(argtypes) == defined[name], (name, len(argtypes), defined[name])


def test_graph_call_structure_helpers_roundtrip():
    from genrl_b200.graphs import _flatten, _rebuild
    a, b, c = torch.zeros(2), torch.ones(3), torch.arange(4)
    obj = ({"x": a, "y": {"z": b}}, [c, 5, "s"], 7)
    flat = _flatten(obj, [])
    assert [p for p, _ in flat] == ["/0/x", "/0/y/z", "/1/0"]
    new = [t + 1 for _, t in flat]
    out = _rebuild(obj, iter(new))
    assert torch.equal(out[0]["x"], a + 1) and torch.equal(out[0]["y"]["z"], b + 1) and torch.equal(out[1][0], c + 1)
    assert out[1][1:] == [5, "s"] and out[2] == 7


@pytest.mark.skipif(not os.path.isdir("/root/reference/agent"), reason="needs the reference checkout (build container only)")
def test_full_size_state_dict_is_checkpoint_compatible_with_the_reference():
    """at the benchmark configuration (stickman, A = 10, 512-d video embeddings) the mirror's state_dict has the
    reference's keys, shapes and dtypes, the same seeded initial values, and loads the reference's state_dict strictly."""
    from genrl_b200.agent.genrl import GenRLAgent
    from oracle.ref_harness import make_reference_agent
    cfg = genrl_config(device="cpu", precision=32, batch_size=2, batch_length=16, imag_horizon=3)
    obs_space, act_spec = make_specs(action_dim=10, clip_dim=512)
    ref = make_reference_agent(cfg, obs_space, act_spec, kind="genrl", seed=1)
    ref_sd = ref.state_dict()
    del ref
    torch.manual_seed(1)
    mine = GenRLAgent(name="genrl", cfg=genrl_config(device="cpu", precision=32, batch_size=2, batch_length=16, imag_horizon=3),
                      obs_space=obs_space, act_spec=act_spec)
    sd = mine.state_dict()
    assert list(sd) == list(ref_sd), [k for k in set(sd) ^ set(ref_sd)][:10]
    n = 0
    for k, v in ref_sd.items():
        assert sd[k].shape == v.shape and sd[k].dtype == v.dtype, k
        assert torch.equal(sd[k], v), f"{k}: seeded initial values differ"
        n += v.numel()
    assert n > 59_000_000                                     # WM 43.3 M (incl. connector) + actor / critic / target critic = 59.6 M
    shuffled = {k: v + 1 for k, v in ref_sd.items() if v.is_floating_point()}
    missing = mine.load_state_dict({**ref_sd, **shuffled}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert torch.equal(mine.state_dict()["wm.rssm._cell._layer.weight"], ref_sd["wm.rssm._cell._layer.weight"] + 1)
