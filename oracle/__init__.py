"""CPU oracle for the factor-graph LM hot path (TEST INFRASTRUCTURE, not product code).

This package is a numpy fp64 restatement of the arithmetic that the reference
(rising-turtle/graph_slam @ 237ad2bf) delegates to GTSAM 4.0 / g2o behind
CGraphGT::optimizeGraphBatch (gtsam/gtsam_graph.cpp:1784-1788),
CGraphG2O::optimizeGraph (g2o/g2o_graph.cpp:241-252) and CImuBase::predictNext
(gtsam/imu_base.cpp:72-87).  GTSAM and g2o are NOT vendored in the reference
(.SUBMODULES.json:8) and are not installed here, so this is a restatement of
their published algorithms, anchored on the reference's
call sites.

Pinning status:
  * OrientedPlane3 / OrientedPlane3Factor: PINNED against the reference's vendored
    known-answer tests gtsam/test/testOrientedPlane3.cpp:61-91,143-164 and
    gtsam/test/testOrientedPlane3Factor.cpp:37-126 (tests/test_oracle_kat.py).
  * BetweenFactor<Pose3>, PriorFactor, GenericProjectionFactor/Cal3DS2,
    CombinedImuFactor / preintegration, LevenbergMarquardt, g2o EdgeSE3:
    PARITY UNPINNED -- the reference holds no golden vectors for them; they are
    validated by central-difference Jacobian checks and closed-form minimisers
    (tests/test_oracle_jacobians.py) and against implementations that share nothing
    with this package (tests/test_oracle_independent.py: scipy expm / logm / Rotation,
    a scalar camera model, brute-force IMU integration, scipy least_squares).

oracle/cpu_lm.cpp (+ cpu_baseline.py) is a TIMING baseline: an OpenMP C++ port of one LM
iteration of the BA + IMU graph, checked against this numpy oracle in tests/test_cpu_baseline.py.
It is not the parity checker (it shares the per-factor math headers with the product).

Only tests/, __graft_entry__ (build of the checker, smoke()) and bench.py's cpu_baseline /
--impl reference legs may import this package.  The product (graph_slam_b200) never does.
"""
