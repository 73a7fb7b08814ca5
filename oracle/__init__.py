"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's geometric hot path, used as the parity checker by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
Nothing under b200pc (the product) imports this package.

  strict      ctypes wrapper over oracle/liboracle.so (strict.c): exact fp32 rounding recipe,
              selections by the total order (distance, index).  The parity oracle.
  ref_torch   the reference's own algorithm re-expressed with the same ATen ops it uses on CPU
              (dense [B,N,M] distance matrix + topk/sort, python FPS loop).  This is what the
              CPU baseline times: it is what the reference executes on host cores.
  cpu_backend the same torch port packaged as a `backend` for b200pc.pointinet (CPU timing of the
              PointINet forward, and the same-weights CPU result the GPU test compares with)
  ref_loader  imports the REAL reference from /root/reference (build container only) with
              in-memory stubs for its missing imports; used to pin the two files above and to
              generate tests/golden/.
"""
