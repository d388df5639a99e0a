"""CPU oracle for the RADTTS hot path -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package `radtts_b200` never imports anything from here.
"""
