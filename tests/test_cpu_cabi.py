"""The drop-in boundary is a plain-C ABI: include/hpvg.h must compile as C99 (no C++, no torch types) and a C program
must be able to link libhpvg.so and call it.  No GPU needed: only entry points that do no device work are called."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")
LIBDIR = os.path.join(ROOT, "mindspore-hp-vae-gan_b200", "hpvg")

C_SRC = r"""
#include <stdio.h>
#include <string.h>
#include "hpvg.h"
int main(void) {
  HpvgGenerator g;
  memset(&g, 0, sizeof g);
  g.n_stages = 1; g.nc_im = 3; g.latent_dim = 128;
  g.T[0] = 4; g.H[0] = 24; g.W[0] = 33;
  g.T[1] = 4; g.H[1] = 30; g.W[1] = 41;
  size_t ws = hpvg_generator_sample_workspace(&g, 2);
  int rc = hpvg_generator_sample(&g, NULL, 2, 0, NULL, NULL, NULL, NULL, 0, NULL);   /* null arguments: refused */
  int32_t i0[5], i1[5]; float l0[5], l1[5];
  int rt = hpvg_linear_taps(4, 5, 1, i0, i1, l0, l1);                                /* host-side resize tables */
  printf("version=%d ws=%zu rc=%d taps=%d i0=%d,%d,%d,%d,%d err=[%s]\n", hpvg_version(), ws, rc, rt,
         (int)i0[0], (int)i0[1], (int)i0[2], (int)i0[3], (int)i0[4], hpvg_last_error());
  return 0;
}
"""


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no C compiler")
def test_header_is_c99_and_a_c_program_links_the_library(tmp_path):
    src = tmp_path / "cabi.c"
    src.write_text(C_SRC)
    exe = tmp_path / "cabi"
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", INC, str(src), "-o", str(exe),
           "-L", LIBDIR, "-lhpvg", "-Wl,-rpath," + LIBDIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    line = out.stdout.strip()
    assert "version=" in line and "rc=-2" in line and "taps=0" in line, line
    assert "i0=0,0,1,2,3" in line, line          # align_corners rule for 4 -> 5: r = 0.75 * o
    ws = int(line.split("ws=")[1].split()[0])
    assert ws > 0 and ws % 256 == 0
    assert "generator_sample" in line             # hpvg_last_error() names the refusing entry point
