"""Host-core probe of a GPU box (no GPU work): why is the reference's CPU arm fast on some boxes and minutes per AR step on
others? Prints the CPU model / flags / cgroup quota and the bf16-autocast Linear throughput at the reference's shapes for
several OpenMP team sizes, then times the reference's own DiffHead network evaluation at two team sizes.
usage: python scripts/host_probe.py > gpurun_out/host_probe.txt"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["CUDA_VISIBLE_DEVICES"] = ""


def p(*a):
    print(*a, flush=True)


def main():
    out = subprocess.run("lscpu | egrep -i 'model name|^CPU\\(s\\)|socket|numa node\\(s\\)|thread\\(s\\) per core'; "
                         "lscpu | egrep -o 'amx_bf16|amx_tile|avx512_bf16|avx512f' | sort | uniq | tr '\\n' ' '; echo; "
                         "cat /sys/fs/cgroup/cpu.max 2>/dev/null; free -g | head -2", shell=True, capture_output=True, text=True)
    p(out.stdout)
    import torch
    from oracle import ref_runner as rr
    usable = rr.usable_cpus()
    p(f"cpu_count {os.cpu_count()} usable {usable} torch threads default {torch.get_num_threads()}")
    x = torch.randn(128, 5120)
    lin_bf = torch.nn.Linear(5120, 15360, bias=False).to(torch.bfloat16)
    lin_f32 = torch.nn.Linear(5120, 15360, bias=True)
    best = (None, 1e9)
    for t in sorted({usable, 96, 64, 48, 32, 16, 8}, reverse=True):
        if t > usable:
            continue
        torch.set_num_threads(t)
        res = []
        for lin, xin in ((lin_bf, x.to(torch.bfloat16)), (lin_f32, x)):
            with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
                lin(xin)
                t0 = time.perf_counter()
                for _ in range(3):
                    lin(xin)
                res.append((time.perf_counter() - t0) / 3)
        gf = 2 * 128 * 5120 * 15360 / 1e12
        p(f"threads {t:4d}: bf16-weight Linear {res[0] * 1e3:8.1f} ms ({gf / res[0]:6.2f} TFLOP/s) | fp32-weight under autocast "
          f"{res[1] * 1e3:8.1f} ms ({gf / res[1]:6.2f} TFLOP/s)")
        if res[0] < best[1]:
            best = (t, res[0])
    p(f"best team size for the 128-row Linear: {best[0]}")
    teams = sorted({usable, best[0]}, reverse=True)
    # one Qwen3-14B decoder layer (bf16 weights), 64 rows, through transformers — an AR pass is 40 of these
    from transformers import Qwen3Config, Qwen3Model
    cfg = Qwen3Config(max_position_embeddings=8192, **{**rr.QWEN3_14B, "num_hidden_layers": 1})
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        lm = Qwen3Model(cfg).eval()
    finally:
        torch.set_default_dtype(old)
    xe = torch.randn(1, 64, 5120).to(torch.bfloat16)
    for t in teams:
        torch.set_num_threads(t)
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            times = []
            for _ in range(3):
                t0 = time.perf_counter()
                lm(inputs_embeds=xe, use_cache=False)
                times.append(time.perf_counter() - t0)
        p(f"threads {t:4d}: one Qwen3-14B layer, 64 rows: {[round(v, 3) for v in times]} s  (x40 = one AR pass)")
    del lm
    # the reference's own head network (built on the meta device: its serial initialisers cost a minute)
    ref = rr.rh.import_reference()
    t0 = time.perf_counter()
    with torch.device("meta"):
        head = ref.fh.DiffHead(parallel_num=64, ch_target=32, ch_cond=5120, ch_latent=5120, depth_latent=6, depth_adanln=2,
                               use_swiglu=True).eval()
    head = head.to_empty(device="cpu")
    rr._randomize(head, 2)
    p(f"DiffHead built in {time.perf_counter() - t0:.1f} s")
    xs, ts, cs = torch.randn(2, 64, 32), torch.rand(2), torch.randn(2, 64, 5120)
    for t in teams:
        torch.set_num_threads(t)
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            times = []
            for _ in range(3):
                t0 = time.perf_counter()
                head.net(xs, ts, cs)
                times.append(time.perf_counter() - t0)
                if times[-1] > 20:
                    break
        p(f"threads {t:4d}: TransEncoder.forward (one head evaluation, 128 rows) {[round(v, 2) for v in times]} s")


if __name__ == "__main__":
    main()
