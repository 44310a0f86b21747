def equizip(*iterables):
    """zip that insists on equal lengths (picklable_itertools.extras.equizip); bin/run.py:16-18 pairs up
    `config_changes` path/value arguments with it."""
    lists = [list(it) for it in iterables]
    if len(set(len(x) for x in lists)) > 1:
        raise ValueError("iterables have different lengths")
    return list(zip(*lists))
