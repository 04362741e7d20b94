class ListConfig(list):
    """`omegaconf.listconfig.ListConfig` stand-in (openaimodel.py:102-104 only checks the type and calls list())."""
