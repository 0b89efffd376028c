"""Prints whether the shipped library matches the shipped sources (no rebuild needed on this machine)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ultra_torchdrug_b200 import build
start = time.time()
print("library current:", build.is_current(), "| exists:", os.path.exists(build.LIB_PATH), "| hash file:", os.path.exists(build.HASH_PATH))
from ultra_torchdrug_b200 import _lib
_lib.lib()
print("loaded in %.2f s" % (time.time() - start))
