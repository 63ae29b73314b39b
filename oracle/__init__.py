"""CPU oracle for the FastPitch 1.1 / HiFi-GAN training hot path of DanRuta/xva-trainer.

TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it, and only as the checker or as the timed CPU baseline.
The product path (xva-trainer_b200/) never imports this package and fails loudly without its CUDA library.

Each function is a restatement (not a copy) of the reference algorithm and cites the reference file:line it
follows. The reference has no tests, golden vectors or fixtures of its own for this path, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: tests/golden/make_golden.py imports the unmodified
reference modules from /root/reference in the build container, runs them on seeded inputs and commits the results
under tests/golden/; tests/test_oracle_golden.py checks this package against those files on every CPU test run.
"""
