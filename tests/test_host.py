"""CPU tests of the host side: C ABI surface, pruning bookkeeping (reference tests mirrored), problem builders,
argument validation, multi-process sharding over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy
import pytest

from conftest import GOLDEN, ROOT


def test_abi_library_exports_every_declared_symbol():
    from ppopt_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'ppgpu.h')).read()
    declared = set(re.findall(r'\b(ppgpu_[a-z0-9_]+)\s*\(', header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.ppgpu_version() >= 100
    assert lib.ppgpu_scan_workspace_bytes(1000) >= 8 * 1001


def test_abi_no_torch_types_in_header():
    header = open(os.path.join(ROOT, 'include', 'ppgpu.h')).read()
    assert 'torch' not in header.replace('torch tensor data_ptr', '') and 'at::' not in header


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'ppopt_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.hpp', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                code = '\n'.join(l for l in src.splitlines() if not l.lstrip().startswith(('//', '#', '*', '"""')))
                assert 'ppopt_oracle' not in src and 'twin_binding' not in src and 'libtwin' not in src, f
                assert not re.search(r'(import|include|from|CDLL).*oracle', code), f


# ---- mirrored from the reference's tests/mpqp_solver_tests/test_mpqp_combinatorial.py:10-82
def test_combination_tester_semantics():
    from ppopt_b200.mp_solvers.solver_utils import CombinationTester, generate_children_sets
    t = CombinationTester()
    assert t.check([]) and t.check(set()) and t.check([1, 2, 3])
    t.add_combo([1, 2])
    assert not t.check([1, 2]) and not t.check([0, 1, 2, 5]) and t.check([1, 3]) and t.check({2})
    t.add_combos({(4,), (0, 3)})
    assert not t.check([4]) and not t.check([0, 1, 3]) and t.check([0, 1])
    assert generate_children_sets([], 3) == [[0], [1], [2]]
    assert generate_children_sets([1], 4) == [[1, 2], [1, 3]]
    assert generate_children_sets([0, 1], 5, t) == [[0, 1, 2]][:0] + [c for c in [[0, 1, 2], [0, 1, 3], [0, 1, 4]] if t.check(c)]
    assert generate_children_sets([3], 4) == []
    t2 = CombinationTester()
    t2.add_combo({1, 2})  # the reference ignores plain sets
    assert t2.check([1, 2])


def test_problem_builders_reproduce_reference_inputs():
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import problems
    for name, build in problems.CONFIGS.items():
        g = numpy.load(os.path.join(GOLDEN, name + '.npz'))
        raw = build()
        for k in ('A', 'b', 'c', 'H', 'A_t', 'b_t', 'F'):
            assert numpy.array_equal(numpy.asarray(raw[k], dtype=float), g['raw_' + k]), (name, k)
        if raw['kind'] == 'qp':
            assert numpy.array_equal(raw['Q'], g['raw_Q'])


def test_program_arrays_validation():
    from ppopt_b200 import engine
    from ppopt_b200.mplp_program import MPQP_Program, load_presolved
    p = load_presolved(os.path.join(GOLDEN, 'factory_mpqp.npz'))
    a = engine.program_arrays(p)
    assert a['is_qp'] and a['A'].flags.c_contiguous and a['b'].ndim == 1 and a['n_eq'] == 0
    p.H = numpy.zeros((2, 4))  # transposed all-zero H is accepted (reference fixtures do this)
    assert engine.program_arrays(p)['H'].shape == (4, 2)
    p.H = numpy.ones((2, 4))
    with pytest.raises(ValueError):
        engine.program_arrays(p)
    with pytest.raises(ValueError):
        MPQP_Program(p.A, p.b, p.c, numpy.zeros((4, 2)), p.Q, p.A_t, p.b_t, p.F, equality_indices=[1], presolved=True)
    lp = load_presolved(os.path.join(GOLDEN, 'transport_mplp.npz'))
    assert not engine.program_arrays(lp)['is_qp']


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from ppopt_b200 import solve_mpqp
    from ppopt_b200.mplp_program import load_presolved
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        solve_mpqp(load_presolved(os.path.join(GOLDEN, 'factory_mpqp.npz')))


def test_chunks_tile_the_level_exactly():
    from ppopt_b200.sharding import chunks
    for n in (0, 1, 5, 16, 17, 1000, 70000, 1234567):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                for lo, hi in chunks(n, r, world):
                    assert 0 <= lo < hi <= n
                    seen.append((lo, hi))
            seen.sort()
            assert [x for lo, hi in seen for x in (lo, hi)] == ([0] + [b for _, b in seen[:-1] for b in (b, b)] + [n] if seen else [])
            if n >= 16 * 65536 * world and world > 1:
                assert all(len(chunks(n, r, world)) == 16 for r in range(world))


WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], 'oracle'))
import numpy, torch, torch.distributed as dist
from ppopt_b200 import sharding
import twin_binding
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:' + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
path = os.path.join(sys.argv[1], 'tests', 'golden', 'mpc_n3.npz')
g = numpy.load(path); tw = twin_binding.Twin.from_npz(path)
for lv in range(int(g['n_levels'])):
    cands = g[f'level{lv}_candidates']; n = len(cands)
    status = torch.zeros(n, dtype=torch.uint8)
    for lo, hi in sharding.chunks(n, rank, world):
        status[lo:hi] = torch.from_numpy(tw.eval(tw.masks(cands[lo:hi].tolist())))
    full = sharding.gather_status(status, dist)
    assert numpy.array_equal(full.numpy() & 11, g[f'level{lv}_status'] & 11), (rank, lv)
    idx = torch.nonzero(full & 8).flatten()
    mine = sharding.owned(idx, n, rank, world)
    allm = [None, None]; dist.all_gather_object(allm, mine.tolist())
    assert sorted(allm[0] + allm[1]) == idx.tolist()
    # region payloads: every rank emits its own regions (CPU checker standing in for K5), the raw buffers are gathered
    # through all_gather on padded tensors and must arrive complete, in candidate order, identical on both ranks
    import types
    k_act = cands.shape[1] - tw.n_eq
    eng = types.SimpleNamespace(tdev=torch.device('cpu'), n=tw.n, n_eq=tw.n_eq, t=tw.t, R0=tw.R0)
    bufs = None
    if mine.shape[0]:
        em = [tw.emit(tw.masks([cands[i].tolist()])[0]) for i in mine.tolist()]
        bufs = [torch.from_numpy(numpy.stack([e[1] for e in em])), torch.from_numpy(numpy.stack([e[2] for e in em])),
                torch.from_numpy(numpy.stack([e[3] for e in em])), torch.from_numpy(numpy.stack([e[4] for e in em]))]
    gi, gb = sharding.gather_regions(eng, mine, bufs, k_act, dist)
    assert gi.tolist() == idx.tolist(), (rank, lv)
    if idx.shape[0]:
        ref = [tw.emit(tw.masks([cands[i].tolist()])[0]) for i in idx.tolist()]
        for j in range(4):
            assert numpy.array_equal(gb[j].numpy(), numpy.stack([e[j + 1] for e in ref])), (rank, lv, j)
        st2 = sharding.gather_region_bits(full.clone() & 7, gi, dist, already_global=True, bufs=gb)
        assert numpy.array_equal(st2.numpy() & 8, full.numpy() & 8)
    bits_in = (full & 7).clone()
    bits_in[mine] |= 8
    st3 = sharding.gather_region_bits(bits_in, mine, dist)
    assert numpy.array_equal(st3.numpy() & 8, full.numpy() & 8)
# ---- witnesses under sharding (engine._eval_level): a rank evaluates its chunks packed into one array, scatters status
# bytes AND witness words back, both are all-reduced (disjoint supports); every rank must end with exactly what the two
# ranks produced for their own chunks.  A stand-in engine (CPU checker: status bytes, sequential walk with witnesses) plays
# the kernels; MIN_CHUNK is lowered so that a rank owns several chunks of these small levels.
from ppopt_b200 import engine, _lib
sharding.MIN_CHUNK, sharding.CHUNKS_PER_RANK = 4, 3
W4 = (tw.R0 + 63) // 64
class FakeEngine:
    tdev = torch.device('cpu'); W = tw.W
    def level_eval(self, masks, k_act, status=None, stages=7, lo=0, hi=None, witness=None, parent=None):
        n = masks.shape[0]; hi = n if hi is None else hi
        if status is None: status = torch.zeros((n,), dtype=torch.uint8)
        m = masks[lo:hi].numpy()
        status[lo:hi] = torch.from_numpy(tw.eval(m.view(numpy.uint64)))
        if witness is not None and hi > lo and k_act >= 1:
            cert, _, wit = tw.k2w_witness(m.view(numpy.uint64))
            witness[lo:hi, 0, :] = torch.from_numpy(wit[:, :tw.W].astype(numpy.int64))
        return status
fake = FakeEngine()
for lv in range(int(g['n_levels'])):
    cands = g[f'level{lv}_candidates']; n = len(cands)
    if n < 16:
        continue
    masks = torch.from_numpy(tw.masks(cands.tolist()).view(numpy.int64).reshape(n, tw.W))
    k_act = cands.shape[1] - tw.n_eq
    wit = torch.zeros((n, _lib.WITNESS_SLOTS, tw.W), dtype=torch.int64)
    st = engine._eval_level(fake, masks, k_act, dist, rank, world, wit, None)
    assert numpy.array_equal(st.numpy() & 11, g[f'level{lv}_status'] & 11), (rank, lv)
    want = torch.zeros_like(wit)
    for r in range(world):
        ch = sharding.chunks(n, r, world)
        assert len(ch) > 1
        packed = torch.cat([masks[lo:hi] for lo, hi in ch])
        wp = torch.zeros((packed.shape[0], _lib.WITNESS_SLOTS, tw.W), dtype=torch.int64)
        fake.level_eval(packed, k_act, None, 7, witness=wp)
        off = 0
        for lo, hi in ch:
            want[lo:hi] = wp[off:off + hi - lo]; off += hi - lo
    assert torch.equal(wit, want), (rank, lv)
    has = (wit[:, 0] != 0).any(dim=1)
    assert bool(((masks & ~wit[:, 0]) == 0).all(dim=1)[has].all()) and int(has.sum()) > 0, (rank, lv)
dist.barrier(); dist.destroy_process_group()
print('ok', rank)
'''


def test_level_sharding_world_size_2_gloo(tmp_path):
    """N>1 host path on CPU: each rank evaluates its chunks (CPU checker standing in for the kernels), status bytes
    are all-reduced over gloo, every rank ends with the reference's full status vector; the region payloads of the
    owning ranks are gathered as raw buffers (sharding.gather_regions) and the region bits made global; witnesses written
    by a rank for its packed chunks are scattered back and all-reduced like the status bytes (engine._eval_level)."""
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=300)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def _is_ordered_subset(sub, full):
    j = 0
    for r in sub:
        while j < full.shape[0] and not numpy.array_equal(full[j], r):
            j += 1
        if j == full.shape[0]:
            return False
        j += 1
    return True


def test_presolve_host_steps_reproduce_reference_arrays():
    """base processing (equalities first, parametric rows moved, L2 scaling, implicit / dependent equalities) must give
    bit-identical rows to the reference constructor; redundancy removal (GPU) can only delete rows, so the golden rows
    must appear, in order and bitwise, among the base-processed ones"""
    from conftest import golden_names
    from ppopt_b200 import presolve
    for name in golden_names():
        g = numpy.load(os.path.join(GOLDEN, name + '.npz'))
        A, b, F, A_t, b_t, eq = presolve.base_processing(g['raw_A'], g['raw_b'], g['raw_F'], g['raw_A_t'], g['raw_b_t'],
                                                         g['raw_equality_indices'].tolist())
        assert eq == list(range(int(g['n_eq']))), name
        main = numpy.hstack([A, b, F])
        assert _is_ordered_subset(numpy.hstack([g['A'], g['b'], g['F']]), main), name
        assert _is_ordered_subset(numpy.hstack([g['A_t'], g['b_t']]), numpy.hstack([A_t, b_t])), name
        if not bool(g['raw_post_process']):
            assert numpy.array_equal(main, numpy.hstack([g['A'], g['b'], g['F']]))
