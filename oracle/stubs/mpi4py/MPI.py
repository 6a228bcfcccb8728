"""Minimal single-process MPI surface used by the reference."""


class Exception(BaseException):  # noqa: A001 - mirrors mpi4py.MPI.Exception
    pass


class Status:
    def Get_error(self):
        return 0


class Request:
    @staticmethod
    def Waitall(requests):
        if requests:
            raise RuntimeError("single-rank stub: no point-to-point requests expected")


class _Comm:
    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def Barrier(self):
        pass

    def gather(self, x, root=0):
        return [x]

    def bcast(self, x, root=0):
        return x

    def Abort(self, code=1):
        raise SystemExit(f"MPI.Abort({code}) called by the reference")

    def Isend(self, *a, **k):
        raise RuntimeError("single-rank stub: Isend must not be reached")

    def Irecv(self, *a, **k):
        raise RuntimeError("single-rank stub: Irecv must not be reached")


COMM_WORLD = _Comm()


def Finalize():
    pass
