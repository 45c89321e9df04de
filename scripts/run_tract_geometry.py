#!/usr/bin/env python3
"""Stand-in for the reference launcher scripts/run_tract_geometry.py: same role, B200 path.

    python scripts/run_tract_geometry.py [--data DIR] [--out DIR] [--config tract_config.json] [--max-streamlines 100|all]

Defaults follow the reference (max_streamlines=100, comprehensive_tract_geometry_analysis.py:310)."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from src.geometry.comprehensive_tract_geometry_analysis import main  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--data"); ap.add_argument("--out"); ap.add_argument("--config")
    ap.add_argument("--max-streamlines", default="100")
    a = ap.parse_args()
    ms = None if a.max_streamlines in ("all", "none", "None") else int(a.max_streamlines)
    df = main(data_dir=a.data, output_dir=a.out, config_path=a.config, max_streamlines=ms)
    print(f"{len(df)} tract records")
