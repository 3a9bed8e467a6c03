"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU/GPU restatement of the reference algorithm for the Cross-Scale MAE
pretraining hot path (engine_pretrain.train_one_epoch over
models_mae.MAE_ViT_MsLdCeCd).  Nothing in the product package
(`cross-scale-mae_b200/`, import name `csmae_b200`) may import this package;
only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs use it, and only as the checker / the CPU baseline.

Pinning status: the reference ships no tests or golden vectors (SURVEY.md
section 8c), so the restatement is pinned against outputs of the REAL reference
classes executed in the build container (`oracle/ref_loader.py` imports
`/root/reference/models_mae/*.py` verbatim with timm/xformers/pytorch_msssim
shimmed).  `tests/golden/make_golden.py` commits those outputs as fixtures and
`tests/test_oracle_golden.py` checks the restatement against them.  The one
piece that cannot be pinned is timm 0.4.12's Block/PatchEmbed arithmetic
(third-party, not vendored, not installable offline): it is restated in
`oracle/timm_shim.py` from the published timm 0.4.12 semantics -- that part is
"parity unpinned" and says so in DESIGN.md.
"""
