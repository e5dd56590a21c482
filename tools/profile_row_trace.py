"""Role-time trace of the row kernel (library built with -DHVQM4_ROW_TRACE): python tools/profile_row_trace.py [S] [profile]"""
import ctypes, os, sys
os.environ["HVQM4_ROW"] = "1"
sys.argv = [sys.argv[0], sys.argv[1] if len(sys.argv) > 1 else "1024", "1", sys.argv[2] if len(sys.argv) > 2 else "0"]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hvqm4_b200 import api
api.lib().hvqm4_row_trace_dump.restype = None
exec(open(os.path.join(ROOT, "tools", "profile_recon.py")).read().replace("ms = batch.replay(R)", "api.lib().hvqm4_row_trace_dump()\nms = batch.replay(R)\napi.lib().hvqm4_row_trace_dump()"))
