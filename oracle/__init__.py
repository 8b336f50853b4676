"""TEST INFRASTRUCTURE ONLY: CPU checkers for the B200 radiance loop.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package. The product (2019global_b200/) never
does; it fails loudly when its CUDA library is missing.
"""
