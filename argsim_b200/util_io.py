"""file helpers used by the train driver (reference src/util_io.py:7-49)."""
import json
import os


def pform(path, *names, sep=''):
    """`path` (with ~ expanded) joined with the `sep`-joined `names` (src/util_io.py:7-9)."""
    return os.path.join(os.path.expanduser(path), sep.join(str(n) for n in names))


def load_txt(filename, encoding=None):
    """yields the lines of a text file without their newline (src/util_io.py:12-15)."""
    with open(filename, encoding=encoding) as f:
        for line in f:
            yield line[:-1]


def save_txt(filename, lines):
    with open(filename, 'w') as f:
        for line in lines:
            f.write('%s\n' % line)


def load_json(filename):
    with open(filename) as f:
        return json.load(f)
