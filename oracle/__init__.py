"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the VideoMamba mixer hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the CPU baseline -- never as the thing
shipped or measured as the GPU path.  ``videomamba_b200`` never imports this package.

Parity status: the reference (tannerhoalst/VideoMamba) ships NO golden vectors for this
path and its arithmetic lives in two un-vendored wheels that are absent here
(``mamba-ssm==2.3.2.post1``, ``causal-conv1d==1.6.2.post1``, reference
``pyproject.toml:16-17``).  The oracle is therefore pinned as far as is possible in this
container: against the reference's own host code + its in-tree ``_selective_scan_ref``
(``models/videomamba/mamba_simple.py:30-106``) imported live from ``/root/reference``
(``oracle/ref_loader.py``), and against fixtures generated from that run
(``tests/golden/``, generator ``tools/make_golden.py``).  The leaf ops whose source is
not in the tree (causal conv + its single-token update, selective scan, single-step state
update, fused add+norm) are restated from the published reference implementations of those
wheels.  They cannot be pinned against the wheels themselves; they ARE cross-checked on the GPU
box (``tests/test_gpu_leaf_crosscheck.py``, fp32, 2e-5) against an independent copy of the same
upstream kernels: vllm (in the image) carries its own adaptations of ``selective_scan_fwd``
(CUDA), ``selective_state_update``, ``causal_conv1d_fn`` / ``causal_conv1d_update`` and the
Triton norm kernel ("Adapted from state-spaces/mamba v2.2.4").  Status for those leaves:
"pinned to the reference's call sites and to vllm's copies of the upstream kernels, not to the
pinned wheel versions themselves".
"""
