"""Test double with the interface of er3t_b200.solver.Solver, backed by the CPU oracle.  Lets the host-side logic
(mcarats_ng -> scene -> jobs -> mca_out_ng) be exercised without a GPU.  TEST INFRASTRUCTURE ONLY."""

import oracle


class OracleSolver:

    def __init__(self, device=0):
        self.device = device
        self.scene = None
        self.options = None
        self._res = None

    def upload_scene(self, scene, options):
        self.scene, self.options = scene, options

    def run(self, jobs, accumulate=False, stream=None, sync=True):
        self._res = oracle.run(self.scene, self.options, jobs)

    def sync(self):
        pass

    def results(self):
        return self._res

    def stats(self):
        return self._res['stats']

    def close(self):
        pass
