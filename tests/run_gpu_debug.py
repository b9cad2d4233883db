"""Ad-hoc GPU debugging driver (not a test): runs one parity scenario verbosely."""
import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import test_gpu_parity as T

for name in ["test_c1_diffdrive_default_critics_injected_noise", "test_c1_two_iterations_reference_fixture",
             "test_c1_free_running_stays_within_tolerance"]:
    try:
        getattr(T, name)()
        print("PASS", name)
    except Exception:
        print("FAIL", name)
        traceback.print_exc()
for model in ["omni", "ackermann"]:
    for fp in [False, True]:
        try:
            T.test_c3_models_full_critics(model, fp)
            print("PASS c3", model, fp)
        except Exception:
            print("FAIL c3", model, fp)
            traceback.print_exc()
