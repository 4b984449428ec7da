"""Drop-in module path of the reference's llamagen package (its test_llamagen.py does `from llamagen... import ...`).
Only what the SJD hot path needs: the GPT parameter tree (so reference checkpoints load) and the solver; the PyTorch
forward is replaced by the sm_100a engine."""
