"""Stand-in for /app/llama-server on a box without a GPU (test infrastructure, never shipped): the PRODUCT's process
entry point (ggufb200.cli.main: argv contract, key file, HTTP server, scheduler, signal handling) with the CPU oracle
injected where the GPU engine goes.  tests/test_server.py points the reference's start.sh at it."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), HERE]

from fake_engine import OracleEngine  # noqa: E402
from ggufb200 import cli  # noqa: E402


def factory(args):
    return OracleEngine(args.model, n_ctx=args.ctx_size, n_slots=max(1, args.parallel))


if __name__ == "__main__":
    sys.exit(cli.main(engine_factory=factory))
