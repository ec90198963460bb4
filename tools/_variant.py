"""Developer variants of libvlpet.so.  The %globaltimer phase stamps of the K1 kernels (tools/trace_k1.py,
tools/trace_k1_bwd.py) are compiled in only with -DVLPET_TRACE: in the default build they cost the issue-bound epilogues
~30 instructions per chunk.  ``use_trace_lib()`` builds ``vl-pet_b200/build/libvlpet_trace.so`` if needed and makes
``vlpet_b200`` load it (VLPET_LIB); call it BEFORE importing vlpet_b200."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def use_trace_lib() -> str:
    spec = importlib.util.spec_from_file_location("_vlpet_build", os.path.join(ROOT, "vl-pet_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    out = os.path.join(ROOT, "vl-pet_b200", "build", "libvlpet_trace.so")
    b.build(defines=("-DVLPET_TRACE",), out=out)
    os.environ["VLPET_LIB"] = out
    return out
