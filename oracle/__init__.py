"""TEST INFRASTRUCTURE ONLY: CPU restatement of the reference hot path + golden-vector tooling.

Nothing under visper_lm_b200/ imports this package. Allowed importers: tests/, __graft_entry__.smoke(),
bench.py's cpu_baseline / --impl reference legs (as checker or reported baseline only).
"""
