"""CPU oracle for the gym-cellular env step -- TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package (see oracle/gc_oracle.c).
"""
