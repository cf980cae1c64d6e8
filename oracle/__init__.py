"""ORACLE package — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import anything from here, and only as the checker / the CPU baseline.  The product
(``hyperbolic-vae_b200/hvae``) never imports it and has no CPU fallback.

Contents
  geoopt_min/   CPU (torch) restatement of the geoopt surface the reference touches (App. A.1)
  pvae_min/     CPU (torch) restatement of the pvae surface the reference touches  (App. A.2)
  stubs/        import stubs for pytorch_lightning / plotly / imageio / termcolor
  reference_loader.py   imports /root/reference/hyperbolic_vae UNMODIFIED over the shims
                        (authoring container only; /root/reference does not travel)
  ref_port.py   restatement of the reference-OWNED hot path (layers, WrappedNormal, KL, models)
                that does travel; pinned against reference_loader output via tests/golden/.

PARITY STATUS: reference-owned code = pinned (reference files executed verbatim to mint the
golden fixtures); third-party geoopt/pvae arithmetic = "parity unpinned" (packages absent from
/root/reference and not installable offline; restated from their published algorithm).
"""
