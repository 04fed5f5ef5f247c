"""makes tools/replay.py importable from the tests"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from replay import *  # noqa: F401,F403,E402
from replay import quat_pose  # noqa: F401,E402
