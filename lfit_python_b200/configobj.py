"""Minimal stand-in for configobj.ConfigObj (not installed here): flat `key = value  # comment`
files as used by /root/reference/test_data/mcmc_input.dat (read at CVModel.py:729, mcmcfit.py:118).
Values stay strings, keys keep file order."""


class ConfigObj(dict):
    def __init__(self, infile=None):
        super().__init__()
        self.filename = infile
        if infile is None:
            return
        if isinstance(infile, (list, tuple)):
            lines = list(infile)
        else:
            with open(infile, 'r') as f:
                lines = f.readlines()
        for raw in lines:
            line = raw.split('#', 1)[0].strip()
            if not line or '=' not in line:
                continue
            key, value = line.split('=', 1)
            value = value.strip()
            if len(value) >= 2 and value[0] == value[-1] and value[0] in "\"'":
                value = value[1:-1]
            self[key.strip()] = value
