"""CPU oracle for the RTM3D keypoint-heatmap decode path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker or as the
timed CPU baseline -- never as the thing shipped.  The product (``rtm3d_b200``) has no CPU fallback and
raises when its CUDA library is missing.

Modules
  decode_ref   torch restatement of the reference decoder (Tier A + the dormant Tier B branch), op for op,
               so that on the same torch build and device it is bit-identical to ``/root/reference``.
               Pinned: ``tests/test_oracle_pin.py`` compares it with the real reference (imported from
               /root/reference when that exists) and with ``tests/golden/*.npz`` produced by
               ``oracle/make_golden.py`` from the real reference.
  canonical    numpy restatement with the canonical (score desc, flat index asc) order made explicit
               (np.lexsort), independent of ``torch.topk``'s tie behaviour.
  box3d_ref    Tier C (closed-form 3D recovery).  NOT in the reference: PARITY UNPINNED, see its header.
  ref_import   loader of the real reference (only usable where /root/reference is mounted).
"""
