"""Developer experiment: relaxation kernel rate at several dimensions (same load alpha = P/N)."""
import json, os, subprocess, sys
for n, p, b in ((1024, 100, 65536), (2048, 200, 32768), (4096, 400, 16384), (8192, 800, 4096)):
    r = subprocess.run([sys.executable, "bench.py", "--dim", str(n), "--targets", str(p), "--probes", str(b), "--steps", "1",
                        "--warmup", "1", "--no-cpu", "--no-delta", "--no-f32"] + (["--columns", os.environ["COLS"]] if "COLS" in os.environ else []), capture_output=True, text=True, env=dict(os.environ, **{k: v for k, v in [a.split("=") for a in sys.argv[1:]]}))
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        ms = d["stage_ms_per_step"]["relax"]
        fl = d["roofline"]["flips_per_probe"]
        flips_per_s = b * fl / ms * 1e3
        print(f"N={n} P={p} B={b}: relax {ms:.1f} ms, {b/ms*1e3:.0f} probes/s, flips/probe {fl:.0f}, sweeps {d['roofline']['sweeps_per_probe']:.1f}, "
              f"{flips_per_s/1e6:.1f} Mflips/s, column GB/s {flips_per_s*8*n/1e9:.0f}, gemm {d['stage_ms_per_step']['fields_gemm']:.1f} ms")
    except Exception as e:
        print(n, "failed", e, r.stderr[-800:])
