import logging as _logging


def get_logger(name):
    return _logging.getLogger(name)
