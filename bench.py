#!/usr/bin/env python
"""bench.py - candidate active sets / second of the combinatorial mpQP enumeration (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            this repo's CUDA engine (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  the reference algorithm's CPU path on the host cores

Workload (config.workload): BASELINE.json configs[4], the synthetic dense mpQP with 100 constraints / 30 variables /
6 parameters = generate_mpqp(30, 6, 40, seed=0) after the reference's presolve (tests/golden/synthetic_30_6_40_s0.npz),
combinatorial levels 1..L (default L = 5: 78,385,935 candidate active sets; L = 4: 3,940,375; the reference cannot finish more than
L = 3 in any reasonable time).  A "step" is one full pass of the level loop over that program.

  value   device-resident throughput: program constants already in HBM, every kernel of the path runs (K1 rank, K2
          feasibility LP, K3/K4 optimality screen, K5 region emission, K6 next-level generation), region buffers stay
          in HBM.  CUDA events on the launch stream, max over ranks, L2 flushed between steps.
  e2e     the same pass through the public call solve_mpqp(program, mpqp_algorithm.combinatorial) with HOST numpy
          program data: upload, all kernels, download of the region matrices, CriticalRegion objects built.
  roofline  the kernel family with the largest share of the step (K2a feasibility certificates since round 1 v5; K2
          simplex before): useful fp64 flops counted in-kernel (K2a: 2 x steps x R0 x (1+k'); K2: 2 x pivots x live rows
          x columns) / its summed launch durations (CUDA events recorded around each launch inside libppgpu), against
          the fp64 FMA peak measured on this device by a register-resident DFMA loop (MEASURED_PEAKS.json has no fp64
          entry).  The path is fp64-FMA / latency bound, not HBM bound: 8W+2 algorithmic bytes per candidate.
  cpu_baseline  the oracle (numpy/HiGHS port of the reference's per-candidate path) on all host cores over a bounded
          random sample of level-3 candidates of the same program.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    'synthetic_30_6_40_s0': 'synthetic dense mpQP 100 constraints / 30 vars / 6 params (generate_mpqp(30,6,40,seed=0))',
    'mpc_n10': 'explicit MPC double integrator, horizon N=10',
    'ctrl_alloc_n5': 'control allocation, 4 rotors, 5 stacked steps',
}
METRIC = 'candidate active sets/sec'


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------------------------- CPU arm (oracle)
def cpu_level3_candidates(P, oracle, cores):
    """levels 1-2 evaluated in full with the oracle, then the exact level-3 candidate list (pruning as the reference)"""
    tester = oracle.CombinationTester()
    level = oracle.children_of(P.equality_indices, P.m, tester)
    for _ in range(2):
        outs = oracle.evaluate_many(P, level, cores)
        feas = []
        for c, (st, _r) in zip(level, outs):
            if st & 2:
                feas.append(c)
            else:
                tester.add_combo(c)
        nxt = []
        for c in feas:
            nxt.extend(oracle.children_of(c, P.m, tester))
        level = nxt
    return level


def cpu_rate(P, oracle, cands, cores, seconds, rng):
    """candidates/s of the oracle on `cores` processes over a random sample sized for about `seconds` of work"""
    probe = [cands[i] for i in rng.choice(len(cands), size=min(len(cands), 8 * cores), replace=False)]
    t0 = time.perf_counter()
    oracle.evaluate_many(P, probe, cores)
    dt = time.perf_counter() - t0
    rate = len(probe) / max(dt, 1e-6)
    size = int(min(len(cands), max(16 * cores, rate * seconds)))
    sample = [cands[i] for i in rng.choice(len(cands), size=size, replace=False)]
    t0 = time.perf_counter()
    outs = oracle.evaluate_many(P, sample, cores)
    dt = time.perf_counter() - t0
    return size / dt, size, dt, sum(1 for s, _ in outs if s & 2)


def run_reference(args, rank):
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ppopt_oracle as oracle
    path = os.path.join(ROOT, 'tests', 'golden', args.workload + '.npz')
    P = oracle.Program.from_npz(path)
    cores = host_cores()
    rng = numpy.random.default_rng(0)
    t_setup = time.perf_counter()
    cands = cpu_level3_candidates(P, oracle, cores)
    t_setup = time.perf_counter() - t_setup
    per_step = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_rate(P, oracle, cands, cores, per_step, rng)
    tot_n, tot_t = 0, 0.0
    for _ in range(args.steps):
        _r, n, dt, _f = cpu_rate(P, oracle, cands, cores, per_step, rng)
        tot_n += n
        tot_t += dt
    value = tot_n / tot_t
    sample = (f'{tot_n // max(1, args.steps)} random level-3 candidates per step (of {len(cands)}; levels 1-2 + level-3 '
              f'generation done once, untimed, {t_setup:.1f}s), rank/feasibility-LP/optimality-LP/region per candidate, '
              f'HiGHS LP backend')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'candidates/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * tot_t / max(1, args.steps),
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOADS[args.workload] + f', combinatorial levels 1..{args.levels}', 'levels': args.levels,
                   'l2': 'n/a (CPU arm)'},
        'cpu_baseline': {'value': value, 'unit': 'candidates/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'candidates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- GPU arm
class ClockSampler(threading.Thread):
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits',
                                      '-i', str(self.gpu)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(s[1]) for s in self.samples if len(s) > 2 and s[1].replace('.', '', 1).isdigit()]
        mx = [float(s[2]) for s in self.samples if len(s) > 2 and s[2].replace('.', '', 1).isdigit()]
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for s in self.samples:
            for nm, v in zip(names, s[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(numpy.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(self.samples)}


def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from ppopt_b200 import engine, mpqp_algorithm, solve_mpqp
    from ppopt_b200.mplp_program import load_presolved
    path = os.path.join(ROOT, 'tests', 'golden', args.workload + '.npz')
    prog = load_presolved(path)
    L = args.levels
    dev = torch.device('cuda', local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn):
        """one step: L2 flush, barrier, CUDA events around fn on the current stream, max over ranks"""
        flush.zero_()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize(dev)
        ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    eng = engine.Engine(engine.program_arrays(prog), device=local_rank)
    dev_step = lambda: engine.solve(prog, max_levels=L, engine=eng, materialize=False)
    e2e_step = lambda: solve_mpqp_capped(prog, L)

    def solve_mpqp_capped(p, levels):
        # the public entry point is solve_mpqp(program, mpqp_algorithm.combinatorial); the depth cap (the reference has
        # none, bench.py applies the same cap to both arms) is passed through the engine behind it
        if levels >= eng.max_depth:
            return solve_mpqp(p, mpqp_algorithm.combinatorial)
        return engine.solve(p, max_levels=levels)

    for _ in range(args.warmup):
        timed(dev_step)
    eng.counters(reset=True)
    eng.profile(True)
    eng.profile_read(reset=True)
    launches0 = eng.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, units = [], 0
    for _ in range(args.steps):
        ms, sol = timed(dev_step)
        ms_dev.append(ms)
        units = sol.total_candidates
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=2)
    prof = eng.profile_read(reset=True)
    eng.profile(False)
    if os.environ.get('PPGPU_BENCH_VERBOSE'):
        print(f'[rank {rank}] ms/step dev {sum(ms_dev) / len(ms_dev):.1f} kernels ' + str({k: round(v['ms'] / len(ms_dev), 1) for k, v in prof.items() if v['launches']}), file=sys.stderr, flush=True)
    counters = eng.counters()
    launches = eng.launch_count() - launches0
    levels_stats = sol.level_stats
    # ---- end to end through the public API with host buffers
    ms_e2e, h2d, d2h, n_regions = [], 0, 0, 0
    for i in range(max(1, min(args.warmup, 2)) + args.steps):
        ms, s2 = timed(e2e_step)
        if i >= max(1, min(args.warmup, 2)):
            ms_e2e.append(ms)
            h2d, d2h, n_regions = s2.h2d_bytes, s2.d2h_bytes, len(s2.critical_regions)
    fp64_peak = engine.measure_fp64_peak()
    total_launches = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(total_launches)
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    value = units * len(ms_dev) / (sum(ms_dev) * 1e-3)
    e2e_value = units * len(ms_e2e) / (sum(ms_e2e) * 1e-3)
    # dominant kernel family of the timed steps and its in-kernel count of useful fp64 FMAs
    fams = {'k2a_relax': ('k2a_relax_reg_kernel (feasibility certificates)', counters['k2a_work'], counters['k2a_tried']),
            'k2_feas_lp': ('k2_feas_kernel (feasibility simplex)', counters['k2_work'], counters['k2_lps']),
            'k34_kkt_cheb': ('k34_kernel (KKT + Chebyshev screen)', counters['k4_work'], counters['k4_lps'])}
    dom = max(fams, key=lambda f: prof[f]['ms'])
    k2 = prof[dom]
    dom_name, dom_work, dom_units = fams[dom]
    k2_flops = 2.0 * dom_work
    k2_tflops = k2_flops / max(k2['ms'] * 1e-3, 1e-12) / 1e12
    hbm_bytes = (8.0 * eng.W + 2.0) * dom_units  # mask in, status byte in + out
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'r01_k2_traffic.json'))).get('dram_bytes_per_launch')
    except Exception:
        pass
    line = {
        'metric': METRIC, 'value': value, 'unit': 'candidates/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': sum(ms_dev) / len(ms_dev), 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOADS[args.workload] + f', combinatorial levels 1..{L}', 'levels': L,
                   'candidates_per_step': units, 'regions_per_step': n_regions, 'parallelism': f'level-sharded x{world}',
                   'l2': 'flushed between steps (256 MiB write)',
                   'per_level': [[s['candidates'], s['feasible'], s['optimal']] for s in levels_stats]},
        'e2e': {'value': e2e_value, 'unit': 'candidates/s', 'ms_per_step': sum(ms_e2e) / len(ms_e2e),
                'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h), 'call': 'solve_mpqp(program, combinatorial)'},
        'gpu_launches': int(total_launches.item()),
        'roofline': {'kernel': dom_name, 'bound': 'fp64', 'achieved': k2_tflops, 'peak': fp64_peak, 'unit': 'TFLOP/s',
                     'frac': k2_tflops / fp64_peak if fp64_peak else None, 'traffic': traffic,
                     'peak_source': 'measured on this device: register-resident DFMA loop (ppgpu_measure_fp64_peak); '
                                    'MEASURED_PEAKS.json carries no fp64 figure',
                     'launches': k2['launches'], 'avg_launch_ms': k2['ms'] / max(1, k2['launches']),
                     'flops_per_launch': k2_flops / max(1, k2['launches']), 'units': dom_units,
                     'share_of_step': k2['ms'] / max(1e-9, sum(v['ms'] for v in prof.values())),
                     'flop_model': 'useful fp64 FMAs counted in-kernel: K2a steps x R0 x (1+k); K2/K4 pivots x live rows x columns',
                     'k2a': {'tried': counters['k2a_tried'], 'certified': counters['k2a_certified'], 'steps': counters['k2a_steps']},
                     'k2': {'lps': counters['k2_lps'], 'pivots': counters['k2_pivots']},
                     'hbm': {'achieved': hbm_bytes / max(k2['ms'] * 1e-3, 1e-12) / 1e9, 'peak': peaks.get('hbm_gbs'),
                             'unit': 'GB/s', 'note': 'algorithmic bytes are 8W+2 B per candidate: not HBM bound'}},
        'kernels_ms_per_step': {k: v['ms'] / len(ms_dev) for k, v in prof.items() if v['launches']},
        'clocks': sampler.summary(),
        'fp64_peak_tflops': fp64_peak,
    }
    if world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import ppopt_oracle as oracle
        P = oracle.Program.from_npz(path)
        cores = host_cores()
        t0 = time.perf_counter()
        cands = cpu_level3_candidates(P, oracle, cores)
        t_setup = time.perf_counter() - t0
        rate, n, dt, _f = cpu_rate(P, oracle, cands, cores, args.cpu_seconds, numpy.random.default_rng(0))
        line['cpu_baseline'] = {'value': rate, 'unit': 'candidates/s', 'cores': cores, 'kind': 'port',
                                'sample': f'{n} random level-3 candidates of the same program in {dt:.1f}s on {cores} '
                                          f'processes (oracle/ppopt_oracle.py, HiGHS LPs; levels 1-2 + level-3 generation '
                                          f'{t_setup:.1f}s untimed)'}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--levels', type=int, default=5)
    ap.add_argument('--workload', default='synthetic_30_6_40_s0', choices=list(WORKLOADS))
    ap.add_argument('--cpu-seconds', type=float, default=15.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    run_gpu(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
