"""paddlenlp/utils/log.py: the `logger` object the training scripts use (info / warning / error / debug)."""
import logging
import sys

logger = logging.getLogger("paddlenlp")
if not logger.handlers:
    _h = logging.StreamHandler(sys.stderr)
    _h.setFormatter(logging.Formatter("[%(asctime)s] [%(levelname)8s] - %(message)s", "%Y-%m-%d %H:%M:%S"))
    logger.addHandler(_h)
    logger.setLevel(logging.INFO)
    logger.propagate = False
