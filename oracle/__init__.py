"""CPU oracle for the recommendation scoring path.  TEST INFRASTRUCTURE ONLY.

Nothing under robot_ebert_b200/ may import this package; only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs do, and there only as the checker or as the
reported CPU baseline.
"""
