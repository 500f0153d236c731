"""CPU oracle for the Qwen3-ASR hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package, and only as the checker (or the timed CPU baseline).  The product path
(libq3asr.so + the ctypes binding) never does.  PARITY UNPINNED: see mel_oracle.c.
"""
