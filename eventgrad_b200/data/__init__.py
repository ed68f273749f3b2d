from .sources import DataSource, load_source, synthetic_source, mnist_source, cifar10_source  # noqa
from .sampler import ShardSampler, per_rank_batch  # noqa
from .loader import BatchLoader, eval_batches  # noqa
from .augment import decode_augment_torch, draw_augment_params  # noqa
