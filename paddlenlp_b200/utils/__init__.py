"""paddlenlp.utils pieces the data-parallel training scripts import (log, tools.get_env_device, batch_sampler)."""
