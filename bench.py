#!/usr/bin/env python3
"""
bench.py -- canonical k-mers/s through DBG build + NBP compaction at k=301 (BASELINE.json metric).

A "step" = one pass of the hot path over one synthetic genome set:
    spb_load (2-bit pack, both strands) -> spb_build_dbg -> spb_compact
`value`  : k-mer instances / s with the ASCII input already resident in HBM (CUDA events on the
           library's stream, max over ranks); mean over the timed steps, median / min alongside.
`e2e`    : same metric through the host-facing API: pinned HOST ASCII in, H2D inside the timed region,
           and the NBP sequences + origin lists + offsets copied back to pinned host memory.
`roofline`: ONE kernel -- the longest single kernel of the step (timed live with its own CUDA event pair inside the
           library) -- with the bytes it must move by design (`achieved`) and the DRAM bytes it really moved
           (`traffic`, ncu capture of this build committed under profiles/); `pipeline` = SURVEY 8(d) model bytes of the
           whole step / time / (n_gpus x peak).
`result_digest`: device-side checksums of every result buffer; at N > 1 rank 0 also runs the single-GPU path on the same
           input and asserts that the digests are equal (multi-rank parity visible in every SCALE line).
`--impl reference` : the reference's OWN CPU implementation of the path (unmodified lib/Assembler.py + its Cython helpers,
           staged under oracle/_ref by build(); graph-tool replaced by the import shim) at os.cpu_count() threads on a
           bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1], configs[2]
    "c2": dict(n_genomes=10, length=2_000_000, ani=0.97, seed=1, desc="synthetic 10 genomes x 2 Mbp ANI 0.97"),
    "c3": dict(n_genomes=100, length=5_000_000, ani=0.96, seed=2, desc="synthetic 100 genomes x 5 Mbp ANI 0.96"),
    "tiny": dict(n_genomes=4, length=200_000, ani=0.97, seed=5, desc="synthetic 4 genomes x 200 kbp ANI 0.97"),
    # BASELINE configs[3] at the 32-bit position limit of this build: 800 (not 1000) x 5 Mbp = 4.0e9 bases, multi-GPU only
    # (every per-position / per-vertex array is sliced over the ranks); generated on the device, no e2e / CPU legs
    "c4s": dict(n_genomes=800, length=5_000_000, ani=0.96, seed=4, device_gen=True,
                desc="synthetic 800 genomes x 5 Mbp ANI 0.96 (BASELINE config 4 at the 32-bit position limit, 4.0e9 bases; generated on the device)"),
    "c4t": dict(n_genomes=16, length=500_000, ani=0.96, seed=4, device_gen=True, desc="synthetic 16 genomes x 500 kbp ANI 0.96 (generated on the device; test of the c4s plumbing)"),
}
K = 301


def algorithmic_bytes(n_inst, n_v, nbp_bases, k=None):
    """N_inst*(16W+24) + N_v*(8W+32) + B_nbp, W = ceil(2k/64)."""
    W = (2 * (k or K) + 63) // 64
    return n_inst * (16 * W + 24) + n_v * (8 * W + 32) + nbp_bases


# Bytes each single-kernel timer's kernel must move BY DESIGN, as a function of the step's counts:
# what it has to read + write once, not the SURVEY model (which charges a materialised 2k-bit key this design never writes).
KERNEL_BYTES = {
    "k_roll_partition_ms": lambda c: 8 * c["n_inst"] + c["n_bases"] // 2,             # 8 B record out per instance, packed strands in
    "k_partition2_ms": lambda c: 16 * c["n_inst"],                                    # 8 B record in + out
    "k_smem_dedup_ms": lambda c: 8 * c["n_inst"],                                     # 8 B record in
    "k_make_records_roll_ms": lambda c: 12 * c["n_inst"] + c["n_bases"] // 2,
    "k_extract_insert_ms": lambda c: 16 * c["n_inst"] + c["n_bases"] // 2,            # one 8 B slot read + written per instance
    "k_assign_ids_ms": lambda c: 4 * c["n_bases"] + 5 * c["n_vertices"],
    "k_edge_write_ms": lambda c: 8 * c["n_edges"] + c["n_vertices"],
    "k_emit_darts_ms": lambda c: c["nbp_bases"] + 4 * c["nbp_vertices"] + 4 * c["n_vertices"],
    "k_origin_pairs_ms": lambda c: 4 * c["n_bases"] + 4 * c["n_vertices"],
    "k_junction_emit_ms": lambda c: 4 * c["n_bases"],
}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(self.rows))


def make_workload(name, mode):
    from superpang_b200 import synth
    from superpang_b200.assembler import concat_sequences
    w = WORKLOADS[name]
    names, seqs = synth.genome_set(w["n_genomes"], w["length"], w["ani"], seed=w["seed"], mode=mode)
    buf, off = concat_sequences(seqs)
    return w, names, buf, off, synth.n_instances(seqs, K)


def config_for(args, world):
    """The `config` object of the JSON line: defined by the command line only, so both arms print the same one."""
    w = WORKLOADS[args.workload]
    return dict(workload=f"{w['desc']}, k={K}, mode={args.mode}", k=K, mode=args.mode,
                parallelism=("1 GPU" if world == 1 else f"k-mers hash-partitioned over {world} GPUs (NCCL all-to-all of 8-byte records, all-gather of the links); "
                             "graph build + compaction sliced: ids / junction entries / origin pairs by position slice, flags / piece lists / edge rows by "
                             "vertex range, NBP vertex lists + sequences by NBP range; all-gathers of the short lists only; results stay sliced in HBM"),
                l2="flushed between timed iterations (256 MiB write)")


class CpuReference:
    """The reference's CPU path on a bounded sample of the workload (same generator, same ANI / mode / seed; fewer and
    shorter genomes): the UNMODIFIED reference (`kind: "reference"`, oracle/_ref) when build() staged it, else the oracle
    port (`kind: "port"`, 1 core).  The sample FASTA is generated once; every `run()` is one timed pass."""

    def __init__(self, workload, mode, target_inst):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import refrun
        from superpang_b200 import synth
        w = WORKLOADS[workload]
        self.n_g = min(w["n_genomes"], 4)
        self.length = max(2 * K, min(w["length"], target_inst // self.n_g))
        names, seqs = synth.genome_set(self.n_g, self.length, w["ani"], seed=w["seed"], mode=mode)
        self.n_inst = synth.n_instances(seqs, K)
        self.mode = mode
        self.real = refrun.available()
        self.cores = os.cpu_count() if self.real else 1
        if self.real:
            fd, self.fasta = tempfile.mkstemp(suffix=".fa", prefix="spb_ref_sample_")
            os.close(fd)
            synth.write_fasta(self.fasta, names, seqs)
        else:
            self.sd = {n: s.tobytes().decode() for n, s in zip(names, seqs)}

    def run(self):
        """One pass; returns (seconds, parts)."""
        t0 = time.perf_counter()
        if self.real:
            import refrun
            parts = {}
            res = refrun.run_reference(self.fasta, K, threads=self.cores, debug=False, with_origins=True, extras=False, timing=parts)
        else:
            import dbg_oracle
            parts = {}
            res = dbg_oracle.run_oracle(self.sd, K)
        assert res["n_inst"] == self.n_inst
        return time.perf_counter() - t0, parts

    def describe(self, dt, parts):
        what = ("unmodified reference lib/Assembler.py (__init__ :80-256, NBP collection :269-363, origin lookup :1212-1262) from oracle/_ref, "
                "graph-tool replaced by the import shim" if self.real else "oracle/dbg_oracle.py (Python port of the reference path)")
        return dict(value=self.n_inst / dt, unit="k-mers/s", cores=self.cores, kind="reference" if self.real else "port",
                    ref_dir="oracle/_ref" if self.real else None,
                    sample=f"{self.n_g} genomes x {self.length} bp of the same generator ({self.mode}), {self.n_inst} k-mer instances, {dt:.1f} s; {what}",
                    parts_s={k_: round(v, 3) for k_, v in parts.items()}, n_inst_sample=self.n_inst)

    def close(self):
        if self.real:
            try:
                os.remove(self.fasta)
            except OSError:
                pass


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = CpuReference(args.workload, args.mode, args.ref_inst)
    times, parts = [], {}
    for i in range(args.warmup + args.steps):
        dt, parts = ref.run()
        if i >= args.warmup:
            times.append(dt)
    ref.close()
    mean = float(np.mean(times))
    cb = ref.describe(mean, parts)
    v = cb["value"]
    print(json.dumps(dict(
        impl="reference", metric=f"canonical k-mers/s through DBG build+NBP compaction at k={K}", value=v, unit="k-mers/s",
        n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * mean, ms_per_step_median=1e3 * float(np.median(times)),
        ms_per_step_min=1e3 * float(np.min(times)), higher_is_better=True,
        scaling="strong", vs_baseline=None, dtype="u64", data="synthetic",
        config=config_for(args, int(os.environ.get("WORLD_SIZE", "1"))),
        cpu_baseline=cb, e2e=dict(value=v, unit="k-mers/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="snp", choices=["snp", "block"])
    ap.add_argument("--ref-inst", type=int, default=1_000_000, help="k-mer instances of the CPU reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-digest", action="store_true")
    ap.add_argument("--no-k-sweep", action="store_true", help="skip the k = 101 / 501 steps (BASELINE config 5) after the headline run")
    ap.add_argument("--k", type=int, default=301, help="k-mer size (BASELINE config 5 sweeps 101 / 301 / 501)")
    args = ap.parse_args()
    global K
    K = args.k
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from superpang_b200 import _capi, digest

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    device_gen = bool(WORKLOADS[args.workload].get("device_gen"))
    if device_gen:
        from superpang_b200 import synth
        w = WORKLOADS[args.workload]
        assert args.mode == "snp", "device-generated workloads are snp mode"
        dbuf, off = synth.genome_set_device(w["n_genomes"], w["length"], w["ani"], w["seed"], dev)
        n_inst = w["n_genomes"] * (w["length"] - K + 1)
        T = dbuf.numel()
        host, buf, dpad, Tpad = None, None, None, T
        args.no_cpu_baseline = True
    else:
        w, names, buf, off, n_inst = make_workload(args.workload, args.mode)
        host = torch.from_numpy(buf).pin_memory()
        T = host.numel()
        Tpad = (T + world * 16 - 1) // (world * 16) * (world * 16)          # equal parts for the all-gather of the input (e2e, N > 1)
        dpad = torch.empty(Tpad, dtype=torch.uint8, device=dev)
        dbuf = dpad[:T]
        dbuf.copy_(host, non_blocking=True)
    stream = torch.cuda.current_stream()
    eng = _capi.Engine(device=local, stream=stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def build():
        return eng.build_dbg_distributed() if world > 1 else eng.build_dbg()

    def step_resident():
        eng.load(dbuf, off, K)
        build()
        return eng.compact()

    out_bufs = {}

    def step_e2e():
        if world == 1:
            eng.load(host, off, K)          # H2D of the ASCII input inside the library call
        else:
            # every rank copies ITS 1/N of the input over ITS PCIe link, then the parts travel over NVLink (NCCL all-gather)
            part = Tpad // world
            a, b = rank * part, min((rank + 1) * part, T)
            if b > a:
                dpad[a:b].copy_(host[a:b], non_blocking=True)
            _capi._all_gather_inplace(dpad, rank, world, None)
            eng.load(dbuf, off, K)
        build()
        c = eng.compact()
        if world > 1 and getattr(eng, "mg_sliced", False):
            # sliced results: every rank copies the sequences of ITS NBPs into the shared page-locked host buffer that rank 0
            # (the process hosting the reference's single-process tail) reads; offsets and origin lists come from rank 0
            n, nb, no = c["n_nbp"], c["nbp_bases"], c["n_origin_pairs"]
            if out_bufs.get("key") != (n, nb, no):
                out_bufs.update(key=(n, nb, no), shared=_capi.SharedHostBuffer(nb, tag="nbp_seqs"))
                if rank == 0:
                    out_bufs.update(soff=torch.empty(n + 1, dtype=torch.int64).pin_memory(), ooff=torch.empty(n + 1, dtype=torch.int64).pin_memory(),
                                    oidx=torch.empty(max(no, 1), dtype=torch.int32).pin_memory())
            nbytes = eng.shard_to_host("nbp_seqs", out_bufs["shared"].tensor)
            if rank == 0:
                eng._ck(eng.lib.spb_get_nbp_seqs(eng.h, out_bufs["soff"].data_ptr(), None))
                eng._ck(eng.lib.spb_get_nbp_origins(eng.h, out_bufs["ooff"].data_ptr(), out_bufs["oidx"].data_ptr()))
                nbytes += 8 * (n + 1) * 2 + 4 * no
            return c, nbytes
        if rank != 0:
            return c, 0
        n, nb, no = c["n_nbp"], c["nbp_bases"], c["n_origin_pairs"]
        key = (n, nb, no)
        if out_bufs.get("key") != key:
            out_bufs.update(key=key, soff=torch.empty(n + 1, dtype=torch.int64).pin_memory(), seq=torch.empty(max(nb, 1), dtype=torch.uint8).pin_memory(),
                            ooff=torch.empty(n + 1, dtype=torch.int64).pin_memory(), oidx=torch.empty(max(no, 1), dtype=torch.int32).pin_memory())
            eng.set_host_output(out_bufs["seq"])     # from the next step on compact() copies the sequences while it computes the rest
        eng._ck(eng.lib.spb_get_nbp_seqs(eng.h, out_bufs["soff"].data_ptr(), out_bufs["seq"].data_ptr()))
        eng._ck(eng.lib.spb_get_nbp_origins(eng.h, out_bufs["ooff"].data_ptr(), out_bufs["oidx"].data_ptr()))
        return c, 8 * (n + 1) * 2 + nb + 4 * no

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        times, stages, last = [], [], None
        launches0 = eng.lib.spb_launch_count()
        for _ in range(steps):
            flush.fill_(1)                   # flush L2 between timed iterations
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            last = fn()
            e1.record(stream)
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
            st = dict(eng.stage_stats())
            st.update(getattr(eng, "mg_times", {}))
            st.pop("multi_gpu", None)
            st.pop("mg_graph_ms", None)
            stages.append(st)
        if world > 1:
            dist.barrier()
        timed.launches = eng.lib.spb_launch_count() - launches0
        return times, stages, last

    sampler = ClockSampler(local)
    sampler.start()
    times, stages, counts = timed(step_resident, args.steps, args.warmup)
    launches = timed.launches
    if device_gen:                                   # no host copy of this input exists: resident number only
        e2e_times, e2e_last = [float("nan")] * args.steps, (None, 0)
    else:
        e2e_times, _, e2e_last = timed(step_e2e, args.steps, max(3, args.warmup))
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # per-step times: max over ranks of every step, then mean / median / min over the steps
    # hash partition balance: k-mer instances owned per rank (max / mean over the ranks)
    balance = None
    if world > 1 and getattr(eng, "mg_owned_records", None) is not None:
        ow = torch.zeros(world, dtype=torch.float64, device=dev)
        ow[rank] = float(eng.mg_owned_records.item())
        dist.all_reduce(ow)
        balance = dict(owned_instances_max=int(ow.max().item()), owned_instances_mean=float(ow.mean().item()),
                       max_over_mean=float((ow.max() / ow.mean()).item()))
    tt = torch.tensor([times, e2e_times], device=dev, dtype=torch.float64)
    d2h = torch.tensor([float(e2e_last[1])], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(d2h)                         # D2H bytes of all ranks
    times, e2e_times = tt[0].tolist(), tt[1].tolist()
    d2h_bytes = int(d2h.item())
    ms, ms_e2e = float(np.sum(times)), float(np.sum(e2e_times))
    total_inst = n_inst                              # strong scaling: the same genome set, k-mers hash-partitioned over the ranks
    value = total_inst * args.steps / (ms / 1e3)
    e2e_value = total_inst * args.steps / (ms_e2e / 1e3)

    # ---- BASELINE config 5 (k-size sweep on the same genome set), N = 1 only: a few untimed-for-the-headline steps at the other
    # k values, reported next to the headline so that the sweep is measured by the same command
    k_sweep = None
    if world == 1 and args.workload == "c3" and K == 301 and not args.no_k_sweep:
        k_sweep = {}
        for kk in (101, 501):
            ts = []
            for i in range(4):
                flush.fill_(1)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                eng.load(dbuf, off, kk)
                eng.build_dbg()
                cc = eng.compact()
                e1.record(stream)
                torch.cuda.synchronize()
                if i:
                    ts.append(e0.elapsed_time(e1))
            k_sweep[str(kk)] = dict(ms_per_step=float(np.mean(ts)), ms_per_step_min=float(np.min(ts)), n_inst=int(cc["n_inst"]),
                                    value=float(cc["n_inst"] / (np.mean(ts) / 1e3)), n_vertices=int(cc["n_vertices"]), n_nbp=int(cc["n_nbp"]))
        step_resident()                              # back to the headline configuration for the digest below

    # ---- result digest (outside the timed region): device-side checksums of every result buffer of the last step
    dg, dg_single = None, None
    if not args.no_digest:
        sliced = world > 1 and getattr(eng, "mg_sliced", False)
        if sliced:                                   # checksums of the local parts, summed over the ranks (collective)
            dg = digest.result_digest_sliced(eng)
        elif rank == 0:
            dg = digest.result_digest(eng, graph=True)
        if world > 1:
            dist.barrier()
            if rank == 0 and not device_gen:         # the same input through the single-GPU path on this rank
                eng.load(dbuf, off, K)
                eng.build_dbg()
                eng.compact()
                dg_single = digest.result_digest(eng, graph=True)
                assert dg_single == dg, f"multi-GPU result differs from the single-GPU result: {dg} != {dg_single}"
            dist.barrier()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
        st_keys = sorted(stages[0])
        st_mean = {k_: float(np.mean([s.get(k_, 0.0) for s in stages])) for k_ in st_keys}
        ni, nv, nbb = counts["n_inst"], counts["n_vertices"], counts["nbp_bases"]
        # the dominant SINGLE kernel: the longest of the kernels that carry their own event pair inside the library
        kt = {k_: v for k_, v in st_mean.items() if k_ in KERNEL_BYTES and v > 0}
        roof = None
        if kt:
            dom = max(kt, key=kt.get)
            dom_ms = kt[dom]
            per_rank = dom in ("k_roll_partition_ms", "k_partition2_ms", "k_smem_dedup_ms") or getattr(eng, "mg_sliced", False)   # this kernel handles 1/N of the units
            dom_bytes = KERNEL_BYTES[dom](counts) / max(world, 1) if per_rank else KERNEL_BYTES[dom](counts)
            ach = dom_bytes / (dom_ms / 1e3) / 1e9
            traffic = None
            try:                                     # measured DRAM bytes of that kernel per launch (ncu --set full of this build)
                tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                traffic = tr.get(f"{args.workload}:{args.mode}:{world}:{K}", {}).get(dom[:-3])
            except Exception:
                pass
            roof = dict(bound="hbm", kernel=dom[:-3], achieved=ach, peak=peak, unit="GB/s", frac=ach / peak, traffic=traffic,
                        traffic_frac=(traffic / (dom_ms / 1e3) / 1e9 / peak) if traffic else None,
                        kernel_ms=dom_ms, algorithmic_bytes=dom_bytes, peak_source=peak_src,
                        note="achieved = bytes this kernel must move by design (DESIGN.md section 3) / its live CUDA-event duration; "
                             "traffic = dram__bytes_read+write of the same kernel from the committed ncu capture (profiles/)",
                        kernels_ms={k_[:-3]: v for k_, v in sorted(kt.items(), key=lambda x: -x[1])})
        if roof is not None and world > 1 and getattr(eng, "mg_exchange_bytes", None) and st_mean.get("mg_exchange_records_ms"):
            # the one bulk collective of the step: all-to-all of the 8-byte records (+ orientation bitmap all-gather), per rank and direction
            xb, xms = float(eng.mg_exchange_bytes), st_mean["mg_exchange_records_ms"]
            roof["nvlink"] = dict(bytes_per_rank=int(xb), ms=xms, achieved=xb / (xms / 1e3) / 1e9, peak=900.0, unit="GB/s per GPU and direction",
                                  frac=xb / (xms / 1e3) / 1e9 / 900.0, peak_source="NVLink 5 nominal (18 links x 50 GB/s per direction)")
        pipe_bytes = algorithmic_bytes(ni, nv, nbb)
        pipe_ach = pipe_bytes / (ms / args.steps / 1e3) / 1e9
        if roof is not None:
            # achieved = Bytes / time / (n_gpu x peak)
            roof["pipeline"] = dict(model_bytes=pipe_bytes, achieved_per_gpu=pipe_ach / world, frac=pipe_ach / world / peak,
                                    note="SURVEY 8(d) byte model of the whole step / step time / (n_gpus x peak); the model charges a materialised "
                                         "2k-bit key per instance that this design never writes")
        line = dict(
            metric=f"canonical k-mers/s through DBG build+NBP compaction at k={K}", value=value, unit="k-mers/s",
            n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms / args.steps, ms_per_step_median=float(np.median(times)),
            ms_per_step_min=float(np.min(times)), higher_is_better=True,
            scaling="weak" if device_gen else "strong", vs_baseline=None, dtype="u64", data="synthetic",
            config=config_for(args, world),
            counts=dict(n_inst=n_inst, n_vertices=counts["n_vertices"], n_edges=counts["n_edges"], n_nbp=counts["n_nbp"], nbp_bases=counts["nbp_bases"],
                        n_origin_pairs=counts["n_origin_pairs"], mg_protocol=getattr(eng, "mg_protocol", None)),
            e2e=None if device_gen else dict(value=e2e_value, unit="k-mers/s", ms_per_step=ms_e2e / args.steps, ms_per_step_median=float(np.median(e2e_times)),
                                             h2d_bytes_per_step=int(buf.nbytes + off.nbytes), d2h_bytes_per_step=d2h_bytes),
            gpu_launches=int(launches),
            roofline=roof,
            stages_ms=st_mean,
            result_digest=digest.fold(dg) if dg else None,
            result_checksums=dg,
            digest_equals_single_gpu=(dg_single == dg) if dg_single is not None else None,
            clocks=sampler.summary(),
            k_sweep=k_sweep,
            balance=balance,
        )
        if not args.no_cpu_baseline:
            ref = CpuReference(args.workload, args.mode, args.ref_inst)
            dt, parts = ref.run()
            line["cpu_baseline"] = ref.describe(dt, parts)
            ref.close()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
