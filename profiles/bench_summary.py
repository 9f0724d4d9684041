"""Prints the interesting numbers of a bench.py JSON line (stdin)."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    if d.get("impl") == "reference":
        print("reference:", d["value"] / 1e6, "M windows/s", d["cpu_baseline"]["sample"])
        continue
    print(f"value {d['value'] / 1e9:.2f} G windows/s  ms/step {d['ms_per_step']:.2f}  e2e {d['e2e']['value'] / 1e9:.2f} G/s "
          f"({d['e2e']['ms_per_step']:.1f} ms: stage {d['e2e']['ms_stage']:.1f} dev {d['e2e']['ms_device']:.1f} fetch {d['e2e']['ms_fetch']:.1f})")
    if "e2e_text_records" in d:
        f = d["e2e_text_records"]
        print(f"e2e from the doubled text records {f['value'] / 1e9:.2f} G/s ({f['ms_per_step']:.1f} ms: stage {f['ms_stage']:.1f} dev {f['ms_device']:.1f} fetch {f['ms_fetch']:.1f})")
    print("kernel_ms", {k: round(v, 3) for k, v in d["kernel_ms"].items()})
    print("roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), "(survey formula", round(d["roofline"].get("survey_formula", {}).get("frac", -1), 3), ") path", round(d["roofline_path"]["frac"], 3),
          "clocks", d["clocks"])
    c = d["config"]
    print("slow-path fractions", round(c.get("slow_path_fraction_pass1", -1), 3), round(c.get("slow_path_fraction_pass2", -1), 3),
          "partitions", c.get("hash_partitions"), "rounds", c.get("rounds"), "h", round(c["pass2_hit_fraction_h"], 3), "gated", round(c["gated_fraction"], 3))
    if "roofline_atomic" in d:
        ra = d["roofline_atomic"]
        print("atomic roof: pass1", round(ra["k_pass1"]["frac"], 3), "pass2", round(ra["k_pass2"]["frac"], 3))
    if "shard_phase_ms_rank0" in d:
        print("phases (rank 0)", d["shard_phase_ms_rank0"])
    if "parity_check" in d:
        print("parity_check", d["parity_check"])
    if "digest_check" in d:
        print("digest_check", d["digest_check"])
    print("runs", c.get("runs"), "windows/run", c.get("windows_per_run"), "run bytes/window", c.get("run_bytes_per_window"))
    if "shard_phase_ms_rank0" in d:
        print("shard phases (rank 0, wall ms)", d["shard_phase_ms_rank0"])
    if "cpu_baseline" in d:
        print("cpu", d["cpu_baseline"]["value"] / 1e6, "M windows/s")
