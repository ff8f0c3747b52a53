# /* Copyright (C) 2001
#  * Housemarque Oy
#  * http://www.housemarque.com
#  *
#  * Distributed under the Boost Software License, Version 1.0. (See
#  * accompanying file LICENSE_1_0.txt or copy at
#  * http://www.boost.org/LICENSE_1_0.txt)
#  */
#
# /* Revised by Paul Mensonides (2002) */
#
# /* See http://www.boost.org for most recent version. */
#
# ifndef BOOST_PREPROCESSOR_LIST_HPP
# define BOOST_PREPROCESSOR_LIST_HPP
#
# include <libint2/boost/preprocessor/list/adt.hpp>
# include <libint2/boost/preprocessor/list/append.hpp>
# include <libint2/boost/preprocessor/list/at.hpp>
# include <libint2/boost/preprocessor/list/cat.hpp>
# include <libint2/boost/preprocessor/list/enum.hpp>
# include <libint2/boost/preprocessor/list/filter.hpp>
# include <libint2/boost/preprocessor/list/first_n.hpp>
# include <libint2/boost/preprocessor/list/fold_left.hpp>
# include <libint2/boost/preprocessor/list/fold_right.hpp>
# include <libint2/boost/preprocessor/list/for_each.hpp>
# include <libint2/boost/preprocessor/list/for_each_i.hpp>
# include <libint2/boost/preprocessor/list/for_each_product.hpp>
# include <libint2/boost/preprocessor/list/rest_n.hpp>
# include <libint2/boost/preprocessor/list/reverse.hpp>
# include <libint2/boost/preprocessor/list/size.hpp>
# include <libint2/boost/preprocessor/list/to_array.hpp>
# include <libint2/boost/preprocessor/list/to_seq.hpp>
# include <libint2/boost/preprocessor/list/to_tuple.hpp>
# include <libint2/boost/preprocessor/list/transform.hpp>
#
# endif
