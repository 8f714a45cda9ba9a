"""env_gym-style module (`<env_id>_data.env_creator`, reference utils/initialization.py:9-30) that hands the
reference loop the stand-in Pendulum of tests/loop/pendulum.py.  Golden generation only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from loop.pendulum import env_creator  # noqa: E402,F401
