"""CPU oracle for the hash-join hot path — TEST INFRASTRUCTURE ONLY (see oracle_join.c).

Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py may import this."""
from .binding import Oracle, build_oracle, load_reference_check  # noqa: F401
