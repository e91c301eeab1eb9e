"""empty test-only stub: the reference imports decord at module import time only"""
