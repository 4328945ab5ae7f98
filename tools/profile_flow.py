"""cProfile of the MFA-shaped file flow (examples/two_pass_alignment.py) on a synthetic corpus: cumulative and self-time views.
Usage: python tools/profile_flow.py [out_dir] [seconds_of_audio]"""
import cProfile, pstats, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "examples"))
import two_pass_alignment as ex
sys.argv = ["x", sys.argv[1] if len(sys.argv) > 1 else "/tmp/ex", sys.argv[2] if len(sys.argv) > 2 else "1800"]
pr = cProfile.Profile()
pr.enable()
ex.main()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(40)
st.sort_stats("tottime").print_stats(40)
st.sort_stats("tottime").print_callers("astype|cumsum|repeat", 12)
