from .simulator import RingSimulator  # noqa
from .trainer import Trainer  # noqa
