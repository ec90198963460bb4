"""Host side of the drop-in: the frozen BART-base model that calls the PET kernels, the PET-only trainer
(flat bucket + one all-reduce + fused AdamW) and the synthetic multitask batches of BASELINE.json's configs."""
from .config import VLPetConfig, bart_base_vlpet_large, t5_base_vlpet_large, bart_base_vlpet_small, bart_base_vlpet_large_video, tiny_test_config, tiny_t5_test_config, TASKS
from .vlt5 import VLT5
from .vlbart import VLBart, VLBartModel, JointEncoder, BartDecoder, BartEncoderLayer, BartDecoderLayer, Downsample
from .trainer import PetTrainer, GraphedPetTrainer, PetBucket, trainable_names, linear_warmup_lr, weight_initialization
from .data import collate, resize_frames, MultitaskLoader, TaskBatches
from .synthetic import VIDEO_TASKS, task_batch_sizes, make_task_batch, multitask_cycle, shard_batch, batch_nbytes

__all__ = ["VLPetConfig", "bart_base_vlpet_large", "t5_base_vlpet_large", "bart_base_vlpet_small", "bart_base_vlpet_large_video", "VIDEO_TASKS", "tiny_test_config", "tiny_t5_test_config", "TASKS", "VLBart", "VLT5", "VLBartModel", "JointEncoder",
           "BartDecoder", "BartEncoderLayer", "BartDecoderLayer", "Downsample", "PetTrainer", "GraphedPetTrainer", "PetBucket",
           "trainable_names", "linear_warmup_lr", "weight_initialization", "task_batch_sizes", "make_task_batch", "multitask_cycle",
           "shard_batch", "batch_nbytes", "collate", "resize_frames", "MultitaskLoader", "TaskBatches"]
