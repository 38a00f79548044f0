import multiprocess


class ProcessingPool:
    """Same surface the reference uses (mpqp_parrallel_combinatorial.py:6,87,116): Pool(n).map / .clear()."""

    def __init__(self, nodes=None):
        self._pool = multiprocess.Pool(nodes)

    def map(self, f, *iterables):
        if len(iterables) == 1:
            return self._pool.map(f, iterables[0])
        return self._pool.starmap(f, zip(*iterables))

    def clear(self):
        self._pool.close()
        self._pool.join()

    def close(self):
        self._pool.close()

    def join(self):
        self._pool.join()

    def restart(self):
        pass
