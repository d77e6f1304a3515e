"""Host logic of pl_yolo_b200.pipeline that needs no GPU: depth rule, loud failure without CUDA, no oracle import."""
import pytest
import torch

from pl_yolo_b200 import pipeline


def test_default_depth():
    assert pipeline.default_depth(1) == 4 and pipeline.default_depth(32) == 4 and pipeline.default_depth(64) == 4
    assert pipeline.default_depth(65) == 2 and pipeline.default_depth(256) == 2


def test_lanes_argument_check():
    with pytest.raises(ValueError):
        pipeline.Lanes(0)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_path():
    with pytest.raises(RuntimeError, match="CUDA"):
        pipeline.Lanes(2)
    pipe = pipeline.PostprocessPipeline([8, 16, 32], conf_thre=0.01, nms_thre=0.65)
    assert list(pipe.drain()) == []
    heads = [torch.zeros(1, 85, 8, 8), torch.zeros(1, 85, 4, 4), torch.zeros(1, 85, 2, 2)]
    with pytest.raises(RuntimeError):
        pipe.submit(heads)
