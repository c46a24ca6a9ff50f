"""Generates tests/golden/known_answers.json from REFERENCE code compiled as-is (run in the build container only).

 - CasSetup: shaders/ffx_a.h + shaders/ffx_cas.h compiled with A_CPU, exactly as source/taa.hpp:15-19 includes them.
 - halton_2_3<8>: the body of helpers::halton is extracted verbatim from source/helper_functions.hpp:9-17 at run time
   (the surrounding header needs gvk.hpp, which does not exist here) and compiled.
Nothing of the reference is copied into the repository: only the numbers are.
"""
import json
import os
import re
import subprocess
import sys
import tempfile

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "known_answers.json")

src = open(os.path.join(REF, "source", "helper_functions.hpp")).read()
m = re.search(r"static float halton\(int i, int b\) \{.*?\n\t\}", src, re.S)
assert m, "halton not found"
halton_src = m.group(0)

prog = r'''
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include <stdlib.h>
#define A_CPU 1
#include "ffx_a.h"
#include "ffx_cas.h"
%s
int main() {
  const float cases[][3] = {{0.5f,3840,2160},{0.0f,1920,1080},{1.0f,1920,1080},{0.25f,7680,4320},{0.8f,320,180},{2.0f,64,64}};
  printf("{\"cas_setup\":[");
  for (unsigned k = 0; k < sizeof(cases)/sizeof(cases[0]); ++k) {
    varAU4(c0); varAU4(c1);
    CasSetup(c0, c1, cases[k][0], cases[k][1], cases[k][2], cases[k][1], cases[k][2]);
    printf("%%s{\"sharpness\":%%.9g,\"w\":%%d,\"h\":%%d,\"const0\":[%%u,%%u,%%u,%%u],\"const1\":[%%u,%%u,%%u,%%u]}", k ? "," : "",
      cases[k][0], (int)cases[k][1], (int)cases[k][2], c0[0], c0[1], c0[2], c0[3], c1[0], c1[1], c1[2], c1[3]);
  }
  printf("],\"halton_2_3_8_px\":[");
  for (int i = 0; i < 8; ++i) printf("%%s[%%.9g,%%.9g]", i ? "," : "", halton(i + 1, 2) - 0.5f, halton(i + 1, 3) - 0.5f);
  printf("],\"halton_raw\":[");
  for (int i = 1; i <= 32; ++i) printf("%%s[%%.9g,%%.9g]", i > 1 ? "," : "", halton(i, 2), halton(i, 3));
  printf("]}\n");
  return 0;
}
''' % halton_src

with tempfile.TemporaryDirectory() as td:
    c = os.path.join(td, "ka.cpp")
    open(c, "w").write(prog)
    exe = os.path.join(td, "ka")
    subprocess.check_call(["/usr/bin/g++", "-O0", "-ffp-contract=off", "-w", "-I", os.path.join(REF, "shaders"), c, "-o", exe])
    data = json.loads(subprocess.check_output([exe], text=True))
json.dump(data, open(OUT, "w"), indent=1)
print("wrote", OUT, {k: len(v) for k, v in data.items()})
