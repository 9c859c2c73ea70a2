"""``Delphi`` facade (same surface as ``python/repair/api.py:26-63``): ``delphi.repair`` hands out
a fresh :class:`RepairModel`, ``delphi.misc`` a fresh :class:`RepairMisc`."""
from .misc import RepairMisc
from .model import RepairModel


class Delphi():

    __instance = None

    def __new__(cls, *args, **kwargs):
        if cls.__instance is None:
            cls.__instance = super(Delphi, cls).__new__(cls)
        return cls.__instance

    @staticmethod
    def getOrCreate() -> "Delphi":
        return Delphi()

    @property
    def repair(self) -> RepairModel:
        return RepairModel()

    @property
    def misc(self) -> RepairMisc:
        return RepairMisc()

    @staticmethod
    def version() -> str:
        return "0.1.0-b200"

    @staticmethod
    def register_table(name, df):
        from . import catalog
        return catalog.register(name, df)
