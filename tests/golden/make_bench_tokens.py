"""Golden greedy tokens of the benchmark models (CPU oracle, canon mode) -- committed so that bench.py can check the
tokens of the TIMED model on boxes / world sizes where running the oracle is not affordable (tensor-parallel runs).
    python tests/golden/make_bench_tokens.py            # writes tests/golden/bench_tokens.json
The models are the seeded synthetic GGUF files bench.py itself writes (bench.model_path)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_tokens.json")
CASES = [("llama3-8b", "Q4_K_M", 32), ("tinyllama-1.1b", "Q4_K_M", 128)]


def main():
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for model, ftype, n in CASES:
        key = f"{model}/{ftype}/{bench.SEED_DEFAULT:#x}"
        if key in res and len(res[key]["tokens"]) >= n:
            continue
        t0 = time.time()
        path = bench.model_path(model, ftype, bench.SEED_DEFAULT)
        m = O.OracleLlama(path, n_ctx=len(bench.PROMPT) + n + 8, nthreads=os.cpu_count(), mode="canon")
        toks = m.greedy(bench.PROMPT, n)
        res[key] = {"prompt": bench.PROMPT, "tokens": [int(t) for t in toks], "oracle_mode": "canon"}
        print(key, f"{time.time() - t0:.0f}s", toks[:8], flush=True)
        with open(OUT, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
