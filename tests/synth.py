import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plade_b200.synth import *  # noqa: F401,F403
