#!/usr/bin/env python
"""sass_post.py -- post-link step of the library build (cuda-to-sycl-nbody_b200/Makefile).

For every instantiation of the production kernel force_wseg_kernel<R, MINB, MASS> in the built library, and for the
scalar one-body-per-lane kernel force_wscalar_kernel<1, none, unit mass> AUTO uses for small shards:
  1. tools/sass_gen.py   regenerates the unrolled tile body from scratch, proving the new block equivalent to
                         ptxas' block before it is written;
  2. tools/sass_sched.py re-orders ptxas' own instructions of any kernel step 1 declines, verifying every
                         dependence of the block it writes;
  3. otherwise ptxas' code stays as it is.
Either way the arithmetic instructions are ptxas' encodings of the reference's IEEE operations; results are
bit-identical (tests/test_parity_gpu.py).  The library records what was done (nbody_kernel_name() reports it).

    python tools/sass_post.py lib/libnbody_b200.so [--no-gen] [--quiet]
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sass_gen as G  # noqa: E402
import sass_sched as S  # noqa: E402

GEN_MARKER = b"NBODY_SASS_GEN="
GENM_MARKER = b"NBODY_SASS_GENM="
GENS_MARKER = b"NBODY_SASS_GENS="
# the scalar one-body-per-lane kernel AUTO uses for small shards (R = 1, no self-term predicate, unit mass)
SCALAR_KERNEL = "force_wscalar_kernelILi1ELi0ELb0"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("lib")
    ap.add_argument("--kernel", default="force_wseg_kernelILi")
    ap.add_argument("--no-gen", action="store_true", help="only re-order ptxas' code (tools/sass_sched.py)")
    ap.add_argument("--quiet", action="store_true")
    ap.add_argument("--gen-scalar", action="store_true",
                    help="also regenerate the scalar small-shard kernel (off until validated on hardware: GEN_SCALAR=1 in the Makefile)")
    a, rest = ap.parse_known_args()
    gp = argparse.ArgumentParser()
    G.add_options(gp)
    gopt = gp.parse_args(rest)
    sopt = argparse.Namespace(pull_window=3, heur=["qgroup"], yield_mode="hold", keep_order=False, no_reuse=False)
    log = (lambda *x: None) if a.quiet else print
    data = bytearray(open(a.lib, "rb").read())
    names = [n for n in S.function_names(a.lib) if a.kernel in n or SCALAR_KERNEL in n]
    n_gen = n_genm = n_gens = n_sched = 0
    for k in names:
        done = False
        mass = "Lb1EEE" in k
        if SCALAR_KERNEL in k:  # generated or left alone (sass_sched handles the packed kernels only)
            if not a.no_gen and a.gen_scalar:
                try:
                    n_gens += 1 if G.process(a.lib, k, data, gopt, log) else 0
                except (ValueError, AssertionError, SystemExit) as e:
                    log(f"{k}: not generated ({e})")
            continue
        if not a.no_gen:
            try:
                done = G.process(a.lib, k, data, gopt, log)
                if done and mass:
                    n_genm += 1
                elif done:
                    n_gen += 1
            except (ValueError, AssertionError, SystemExit) as e:
                log(f"{k}: not generated ({e})")
        if not done:
            try:
                n_sched += 1 if S.process_kernel(a.lib, k, data, sopt, log) else 0
            except (ValueError, AssertionError, SystemExit) as e:
                log(f"{k}: not scheduled ({e})")
    for marker, count in ((S.MARKER, n_sched), (GEN_MARKER, n_gen), (GENM_MARKER, n_genm), (GENS_MARKER, n_gens)):
        m = data.find(marker)
        if m >= 0:
            data[m + len(marker):m + len(marker) + 2] = b"%02d" % count
    print(f"sass_post: {len(names)} kernels: {n_gen} + {n_genm} (per-body mass) + {n_gens} (scalar small-shard) regenerated, "
          f"{n_sched} re-scheduled, {len(names) - n_gen - n_genm - n_gens - n_sched} left as ptxas wrote them ({a.lib})")
    tmp = tempfile.NamedTemporaryFile(dir=os.path.dirname(os.path.abspath(a.lib)), suffix=".so", delete=False)
    tmp.write(bytes(data))
    tmp.close()
    chk = subprocess.run(["cuobjdump", "-sass", tmp.name], capture_output=True, text=True)
    if chk.returncode != 0 or "error" in chk.stderr.lower():
        os.unlink(tmp.name)
        print("sass_post: patched image does not disassemble; library left untouched")
        return 1
    os.chmod(tmp.name, 0o755)
    os.replace(tmp.name, a.lib)
    return 0


if __name__ == "__main__":
    sys.exit(main())
