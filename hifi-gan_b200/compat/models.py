"""Top-level `models` shim: lets code written against the reference's `src/` layout (`from models import
Generator`, inference.py:11) resolve to hifigan_b200 by putting this directory on PYTHONPATH instead of `src/`."""
from hifigan_b200.models import *  # noqa: F401,F403
from hifigan_b200.models import (LRELU_SLOPE, DiscriminatorP, DiscriminatorS, Generator,  # noqa: F401
                                 MultiPeriodDiscriminator, MultiScaleDiscriminator, ResBlock1, ResBlock2,
                                 discriminator_loss, feature_loss, generator_loss)
