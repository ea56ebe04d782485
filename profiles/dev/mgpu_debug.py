"""Where does a 2-rank train() with CUDA-graph capture stall?  Runs tests/test_multi_gpu.py's worker with a traceback dump."""
import faulthandler, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "colvars-finder_b200")]


def worker(rank, world, port, out):
    faulthandler.dump_traceback_later(int(os.environ.get("DUMP_AFTER", "75")), exit=True, file=sys.stderr)
    from tests.test_multi_gpu import _worker
    _worker(rank, world, port, out)
    print("rank", rank, "finished", flush=True)


if __name__ == "__main__":
    import torch.multiprocessing as mp
    from tests.test_multi_gpu import _free_port
    with tempfile.TemporaryDirectory() as out:
        mp.spawn(worker, args=(2, _free_port(), out), nprocs=2, join=True)
