"""Host-side check of the 16-bit token-row layout of the tensor-core path (csrc/ufo_common.cuh: tok_pos).

The gather's lane j of a sample point owns feature channels 4j..4j+3, channel j of the three frustum features and depth-PE
component j (reference channel order: feat 32 | vol 24 | sim 16 | PE 8, code1/ray_transformer.py:258-288).  tok_pos must be a
permutation of 0..79 that makes those eight values one contiguous, 16-byte aligned piece of the row, and keeps the 16 pre_sim_mlp
outputs as the last two 16-byte pieces.  The function is compiled for the host with nvcc (no GPU needed)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r"""
#include <cstdio>
#include "ufo_common.cuh"
int main() {
  for (int c = 0; c < 80; ++c) printf("%d\n", ufo::tok_pos(c));
  return 0;
}
"""


def _tok_pos(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    src = tmp_path / "tokpos.cu"
    src.write_text(SRC)
    exe = tmp_path / "tokpos"
    subprocess.run([nvcc, "-std=c++17", "-I", os.path.join(ROOT, "uforecon_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
                    "-o", str(exe), str(src)], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    return [int(v) for v in out]


def test_tok_pos_is_a_lane_major_permutation(tmp_path):
    pos = _tok_pos(tmp_path)
    assert sorted(pos) == list(range(80)), "tok_pos must be a permutation of the 80 token channels"
    for j in range(8):
        mine = [pos[4 * j + i] for i in range(4)]                 # feature quad of lane j
        mine += [pos[32 + 8 * s + j] for s in range(3)]           # channel j of the three frustum features
        mine += [pos[72 + j]]                                     # depth-PE component j
        assert mine == list(range(8 * j, 8 * j + 8)), f"lane {j}: its eight values must be one aligned 16-byte piece, in this order"
    assert [pos[56 + i] for i in range(16)] == list(range(64, 80)), "pre_sim_mlp outputs: the last two 16-byte pieces"
