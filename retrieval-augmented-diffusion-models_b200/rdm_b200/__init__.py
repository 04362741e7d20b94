"""rdm_b200: B200-native implementation behind the reference's ``rdm.*`` import paths.

``_lib`` binds librdm_b200.so (hand-written sm_100a CUDA, C ABI in include/rdm_b200.h);
the sibling ``rdm/`` package mirrors the reference's class paths on top of it.
"""
