"""ORACLE - TEST INFRASTRUCTURE ONLY.  Builds the native checkers into oracle/_ref/ (git-ignored, travels to the GPU box):

  libfastutils_port.so  the C restatement oracle/native/fast_utils_port.c                       (always)
  libfastutils_ref.so   the reference's own find_peaks.cpp + assign.cpp, compiled where they lie   (only when
                        /root/reference is present: this container, not the GPU box)

Plain gcc/g++ on the few source files; the reference's build system (torch cpp_extension, plugins.cpp) is not used.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "_ref")
REF_PARSE = "/root/reference/nano_demo/fast_utils/parse"


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(verbose=False):
    os.makedirs(OUT, exist_ok=True)
    built = {}
    port_src = [os.path.join(HERE, "fast_utils_port.c")]
    port = os.path.join(OUT, "libfastutils_port.so")
    if _newer(port, port_src):
        # -ffp-contract=off: the restatement must not fuse d*100 - val (the reference build targets plain x86-64)
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", port] + port_src + ["-lm"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    built["port"] = port
    ref_src = [os.path.join(REF_PARSE, "find_peaks.cpp"), os.path.join(REF_PARSE, "assign.cpp")]
    if all(os.path.exists(s) for s in ref_src):
        ref = os.path.join(OUT, "libfastutils_ref.so")
        srcs = [os.path.join(HERE, "fast_utils_refwrap.cpp")] + ref_src
        if _newer(ref, srcs):
            cmd = ["g++", "-O2", "-fPIC", "-shared", "-I", REF_PARSE, "-o", ref] + srcs
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
        built["ref"] = ref
    elif os.path.exists(os.path.join(OUT, "libfastutils_ref.so")):
        built["ref"] = os.path.join(OUT, "libfastutils_ref.so")      # prebuilt copy that travelled with the snapshot
    return built


if __name__ == "__main__":
    print(build(verbose=True))
