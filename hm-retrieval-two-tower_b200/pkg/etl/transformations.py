"""DataFrame helpers of the reference's ETL (pkg/etl/transformations.py:9-94) that the in-scope runners call: CSV loading with
pandas' dtype inference, the inclusive date-range filter, CSV saving.  Host-side pandas; no numerics."""
from __future__ import annotations

import logging
import os
from typing import Tuple

logger = logging.getLogger(__name__)


def date_filter(df, df_name: str, date_col: str, date_range: Tuple[str, str]):
    """Rows with ``date_range[0] <= df[date_col] <= date_range[1]`` (both ends inclusive; ISO ``YYYY-MM-DD`` strings compare in
    calendar order, which is what the reference relies on)."""
    start, end = date_range
    logger.info(f"Creating df {df_name} from: {start} to: {end}")
    col = df[date_col]
    kept = df[(col >= start) & (col <= end)]
    logger.info(f"{df_name}: {len(kept)} of {len(df)} rows")
    return kept


def load_dataframe(path: str, df_name: str):
    """``pd.read_csv(path)`` with default dtype inference: an all-digit id column comes back as integers, so "0108775015" is
    108775015 on every path that later stringifies it (vocabularies, logQ table, TFRecords, popularity index)."""
    import pandas as pd

    logger.info(f"Loading {df_name} from {path}")
    df = pd.read_csv(path)
    logger.info(f"{df_name}: {len(df)} rows")
    return df


def save_dataframe(df, df_name: str, date_col: str, path: str) -> None:
    """Write ``df`` as CSV without the index, creating the directory; logs the date span being written."""
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    if len(df):
        logger.info(f"Saving {df_name} ({len(df)} rows, {df[date_col].min()} .. {df[date_col].max()}) to {path}")
    else:
        logger.info(f"Saving empty {df_name} to {path}")
    df.to_csv(path, index=False)
