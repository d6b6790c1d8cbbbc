"""Drop-in shim: ``import DINCAE`` resolves to the B200-native implementation, so
``run_DINCAE.py`` of the reference runs unchanged (DINCAE/__init__.py:47 public names)."""
from dincae_b200 import *                                        # noqa: F401,F403
from dincae_b200 import (FrameSource, NEAREST_NEIGHBOR, data_generator_list, identity,
                         reconstruct_gridded_files, save_checkpoint, savesample, sinv)  # noqa: F401
from dincae_b200 import __all__, __version__                     # noqa: F401
