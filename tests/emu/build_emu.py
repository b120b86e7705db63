"""Builds the host-emulation library of the kernels for one model (test infrastructure, CPU only)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(ROOT, "learning-from-sparse-demonstrations_b200", "csrc")


def build(model_header, out_so, tsan=False, extra=(), ns="cpdp_emu"):
    cmd = ["g++", "-std=c++20", "-O1" if tsan else "-O2", "-g", "-fPIC", "-shared", "-pthread",
           "-fvisibility=hidden", "-fno-gnu-unique", "-DCPDP_NS=%s" % ns,
           "-I", CSRC, "-DCPDP_MODEL_HEADER_PORT=\"cpdp_port.h\"", "-DCPDP_MODEL_HEADER=\"%s\"" % model_header,
           os.path.join(ROOT, "tests", "emu", "emu_lib.cpp"), "-o", out_so]
    if tsan:
        cmd[1:1] = ["-fsanitize=thread"]
    cmd += list(extra)
    subprocess.check_call(cmd)
    return out_so


if __name__ == "__main__":
    build(sys.argv[1], sys.argv[2], tsan="--tsan" in sys.argv)
