"""test-only stub: the reference imports decord and sets its bridge at module import time only"""


class bridge:  # noqa: N801
    @staticmethod
    def set_bridge(name):
        return None
