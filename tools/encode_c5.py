"""BASELINE.json configs[4] (the --gen-specgram / validation encode path): thin wrapper around `bench.py --workload encode`.

  python tools/encode_c5.py
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/encode_c5.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

if __name__ == "__main__":
    sys.argv = [sys.argv[0], "--workload", "encode"] + sys.argv[1:]
    bench.main()
