timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 500 -k "full_grid_all_phi or c3_recipe_full" 2>&1 | tail -15
