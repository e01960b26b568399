"""TEST INFRASTRUCTURE ONLY -- see the package docstring."""


class mOTU:
    def __init__(self, name="mOTUpan_core_prediction", **kwargs):
        self.name = name

    def get_stats(self):
        return {self.name: {"core": None, "aux_genome": [], "singleton_cogs": []}}
