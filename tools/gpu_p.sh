#!/bin/bash
timeout 400 python -m pytest tests/test_k1_single_level_gpu.py tests/test_checkpoint_gpu.py tests/test_virtual_ranks_gpu.py tests/test_launch_variants_gpu.py tests/test_graph_replay_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()"
