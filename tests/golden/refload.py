"""Import the UNMODIFIED reference module from /root/reference (this container only).

Nothing under tests that runs on the GPU box may call this: /root/reference does not exist
there.  It is used by make_golden.py (fixture generation) and by the optional live-parity
tests, which skip when the reference tree is absent.
"""
import importlib.util
import logging
import os
import sys

REF_ROOT = "/root/reference"
REF_FILE = os.path.join(REF_ROOT, "src", "ancestral_reconstruction.py")


def reference_available():
    return os.path.exists(REF_FILE)


def load_reference():
    if not reference_available():
        raise FileNotFoundError(REF_FILE)
    stub = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ete3_stub")
    if stub not in sys.path:
        sys.path.insert(0, stub)
    name = "_reference_ancestral_reconstruction"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    logging.getLogger().setLevel(logging.ERROR)  # the reference logs at INFO on import
    return mod
