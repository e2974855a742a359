"""TEST INFRASTRUCTURE ONLY: ctypes binding to the CPU oracle (oracle/_ref/libcfref.so).

The oracle is the reference's own block encoders compiled from /root/reference by
oracle/Makefile, driven by oracle/cfref.cpp (our restatement of Cuttlefish's Converter glue).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (cuttlefish_b200) never does.
"""
from .cfref import *  # noqa: F401,F403
