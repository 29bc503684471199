"""Random ego networks in the schema of TweetRecommender/SQLiteAdapter.cs:27-125, for the tests of the callers' rows (N1-N4):
as table rows (for the reference's DataLoader compiled from its sources, oracle/ref.py) and as a *.sqlite file (for the
product's ingest).  Valid for Experiment.cs:72-74 by default: >= 50 likes of the ego user, >= 50 friends."""
import os
import sqlite3

SCHEMA = """CREATE TABLE follow(source INTEGER, target INTEGER); CREATE TABLE tweet(id INTEGER, author INTEGER);
CREATE TABLE retweet(user INTEGER, tweet INTEGER); CREATE TABLE quote(user INTEGER, tweet INTEGER);
CREATE TABLE favorite(user INTEGER, tweet INTEGER); CREATE TABLE mention(source INTEGER, target INTEGER);"""


def random_tables(rng, n_friends=55, n_nonfriend_followees=6, n_third=25, n_tweets=260, ego=1000, ego_likes=70):
    """`rng`: random.Random.  The ego follows `n_friends` users who follow back (the members) and a few who do not; members
    follow members and third-party users; likes arrive through the three tables, some twice; tweets have member, third-party
    and unknown authors; mentions between the first 20 members.  Duplicate rows on purpose."""
    members = [ego] + [ego + 1 + i for i in range(n_friends)]
    nonfriends = [ego + 500 + i for i in range(n_nonfriend_followees)]
    third = [ego + 900 + i for i in range(n_third)]
    follow = [(ego, m) for m in rng.sample(members[1:] + nonfriends, n_friends + n_nonfriend_followees)]
    for m in members[1:]:
        outs = [ego] + rng.sample(members[1:] + third, rng.randrange(2, 9))
        rng.shuffle(outs)
        follow += [(m, t) for t in outs if t != m]
    follow.append(follow[rng.randrange(len(follow))])
    tweets = [5_000_000 + 7 * i for i in range(n_tweets)]
    tweet = [(t, rng.choice(members + third + [999_999])) for t in tweets]
    rng.shuffle(tweet)
    likes = {"retweet": [], "quote": [], "favorite": []}
    for m in members:
        for t in rng.sample(tweets, ego_likes if m == ego else rng.randrange(0, 14)):
            likes[rng.choice(list(likes))].append((m, t))
            if rng.random() < 0.1:
                likes[rng.choice(list(likes))].append((m, t))
    mention = []
    for _ in range(400):
        a, b = rng.sample(members[:20], 2)
        mention += [(a, b)] * rng.randrange(1, 4)
    rng.shuffle(mention)
    return dict(follow=follow, tweet=tweet, mention=mention, **likes)


def add_nasty_rows(tables, rng, ego=1000):
    """Rows a crawler can produce and DataLoader has an answer for: self-follows, followers the ego does not follow, mentions of
    oneself and of / by non-members, a duplicate tweet row, likes of a tweet that is not in the tweet table; all tables shuffled."""
    m1, m2 = ego + 1, ego + 2
    tables["follow"] += [(ego, ego), (m1, m1), (7777, ego), (m2, 7777)]
    tables["mention"] += [(ego, ego)] * 3 + [(ego, 7777)] * 2 + [(7777, m1)] * 2 + [(m1, m2)] * 5
    tables["tweet"] += [tables["tweet"][0], (4242, ego), (4243, m1)]
    tables["favorite"] += [(ego, 4242), (m1, 4242), (m1, 999999), (ego, 4243)]
    for k in tables:
        rng.shuffle(tables[k])
    return tables


def write_sqlite(path, tables):
    if os.path.exists(path):
        os.remove(path)
    c = sqlite3.connect(path)
    c.executescript(SCHEMA)
    for name, rows in tables.items():
        c.executemany(f"INSERT INTO {name} VALUES (?, ?)", rows)
    c.commit()
    c.close()
    return path
