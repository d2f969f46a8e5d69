"""Logging in the reference's style (src/logUtils.py:145-147: one logger named "BPReveal",
``LEVEL : time :file:line : message``) -- only what the hot-path programs call."""
from __future__ import annotations

import logging

DEBUG, INFO, WARNING, ERROR, CRITICAL = (logging.DEBUG, logging.INFO, logging.WARNING,
                                         logging.ERROR, logging.CRITICAL)
_logger = logging.getLogger("BPReveal")
if not _logger.handlers:
    _h = logging.StreamHandler()
    _h.setFormatter(logging.Formatter(
        "%(levelname)s : %(asctime)s :%(filename)s:%(lineno)d : %(message)s"))
    _logger.addHandler(_h)
    _logger.setLevel(logging.WARNING)
    _logger.propagate = False


def setVerbosity(userLevel):
    """``userLevel``: "DEBUG" | "INFO" | "WARNING" | "ERROR" | "CRITICAL" (base.schema:23-26)."""
    _logger.setLevel(getattr(logging, userLevel) if isinstance(userLevel, str) else userLevel)


def getVerbosity():
    return _logger.level


def debug(msg, *a):
    _logger.debug(msg, *a, stacklevel=2)


def info(msg, *a):
    _logger.info(msg, *a, stacklevel=2)


def warning(msg, *a):
    _logger.warning(msg, *a, stacklevel=2)


def error(msg, *a):
    _logger.error(msg, *a, stacklevel=2)


def critical(msg, *a):
    _logger.critical(msg, *a, stacklevel=2)
