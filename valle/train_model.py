"""``python -m valle.train_model -c cfg.json -m ValleAR`` -- the reference's entry point (valle/train_model.py:38-44), served by
valle2_b200.train_model."""
from valle2_b200.train_model import batches, fit, main, synthetic_items, train  # noqa: F401

if __name__ == '__main__':
    main()
