"""Logger the host layer writes its timing lines to (mirrors modal.config.logger usage in the reference)."""
import logging
import os

logger = logging.getLogger("modal_client_b200")
_level = os.environ.get("MODAL_LOGLEVEL")
if _level:
    logger.setLevel(_level.upper())
