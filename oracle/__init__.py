"""oracle/ -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package, and only as the checker / CPU baseline -- never as the thing shipped.  The product
package (nerf-loam_b200/) must not import it (tests/test_no_oracle_in_product.py enforces that).
"""
