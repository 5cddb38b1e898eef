"""The drop-in boundary, compile-checked: cuda-phdslam_b200/shim/phdfilter_b200.cpp is the file a maintainer of the
reference compiles INSTEAD of src/phdfilter.cu.  Here it is compiled against the reference's own headers
(src/slamtypes.h, src/phdfilter.h) and linked, together with a stub of main.cpp's globals and call sites
(src/main.cpp:1200,1251-1253,1271,1464-1465), against libphdslam.so: every symbol resolves.  (Running it needs a GPU.)"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src"
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "phdfilter.h")), reason="reference headers not present (GPU box)")

MAIN_STUB = r"""
class MotionModel;
#include "slamtypes.h"
#include "phdfilter.h"
SlamConfig config;                       /* src/main.cpp:80 */
void recoverSlamStateDevice(SynthSLAM&, ConstantVelocityState&, vector<REAL>&);
int main(int argc, char**) {
  if (argc < 100) return 0;              /* link check only: creating a handle needs a GPU */
  initRandomNumberGenerators();          /* src/main.cpp:1464 */
  setDeviceConfig(config);               /* :1465 */
  SynthSLAM particles(config.n_particles);
  AckermanControl u; u.alpha = 0; u.v_encoder = 1;
  phdPredict(particles, u);              /* :1253 */
  phdPredict(particles);                 /* :1251 */
  measurementSet Z(3);
  SynthSLAM pre = phdUpdateSynth(particles, Z);   /* :1271 */
  ConstantVelocityState e; vector<REAL> cn;
  recoverSlamStateDevice(particles, e, cn);
  return (int)pre.n_particles;
}
"""


@needs_ref
def test_reference_side_shim_compiles_and_links(tmp_path):
    lib_dir = os.path.join(ROOT, "cuda-phdslam_b200")
    assert os.path.exists(os.path.join(lib_dir, "libphdslam.so")), "build first: python __graft_entry__.py"
    shim = os.path.join(lib_dir, "shim", "phdfilter_b200.cpp")
    stub = tmp_path / "main_stub.cpp"
    stub.write_text(MAIN_STUB)
    flags = ["-std=c++14", "-fpermissive", "-w", "-I", REF, "-I", os.path.join(ROOT, "include")]
    obj = str(tmp_path / "shim.o")
    subprocess.check_call(["g++"] + flags + ["-c", shim, "-o", obj])
    exe = str(tmp_path / "dropin")
    subprocess.check_call(["g++"] + flags + [str(stub), obj, "-L", lib_dir, "-lphdslam", "-Wl,-rpath," + lib_dir, "-o", exe])
    # the shim defines the reference's own symbols (C++ linkage, the reference's types)
    syms = subprocess.check_output(["nm", "-C", "--defined-only", obj], text=True)
    for name in ("initRandomNumberGenerators()", "setDeviceConfig(SlamConfig const&)", "phdPredict(SynthSLAM&, ...)",
                 "phdUpdateSynth(SynthSLAM&,"):
        assert name in syms, name
    assert subprocess.call([exe]) == 0      # loads libphdslam.so (and libcudart) and exits before touching a device
